// Shared helpers for the vivae_b200 CUDA kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/vivae_b200.h"

namespace vvb {

// ---------------------------------------------------------------- error plumbing
void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;  // kernels launched by this library
inline void count_launch(int64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define VVB_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::vvb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,     \
                       __LINE__);                                                            \
      return VVB_ECUDA;                                                                      \
    }                                                                                        \
  } while (0)

#define VVB_CHECK_LAUNCH()                                                                   \
  do {                                                                                       \
    ::vvb::count_launch();                                                                   \
    VVB_CUDA(cudaGetLastError());                                                            \
  } while (0)

#define VVB_REQUIRE(cond, ...)                                                               \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      ::vvb::set_error(__VA_ARGS__);                                                         \
      return VVB_EINVAL;                                                                     \
    }                                                                                        \
  } while (0)

// a valid request outside the implemented envelope (Python: NotImplementedError, never the reference's ValueError)
#define VVB_SUPPORTED(cond, ...)                                                             \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      ::vvb::set_error(__VA_ARGS__);                                                         \
      return VVB_ENOTSUP;                                                                    \
    }                                                                                        \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// carve sub-buffers out of one workspace allocation (256-byte aligned)
struct Carver {
  char *base;
  size_t off = 0;
  explicit Carver(void *p) : base(static_cast<char *>(p)) {}
  template <typename T>
  T *take(size_t count) {
    off = align_up(off, 256);
    T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
  size_t used() const { return align_up(off, 256); }
};

int sm_count();

#ifdef __CUDACC__
// ---------------------------------------------------------------- device helpers
constexpr unsigned FULL = 0xffffffffu;

// monotone map float -> uint32 (a < b  <=>  ord(a) < ord(b), -0 < +0)
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(b);
}

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed(const int *p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void st_release(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- in-register bitonic sort of 32*E 64-bit keys spread over a warp ------------------
// element p = e*32 + lane holds key[e]; result ascending in p.
template <int V> struct Log2 { static constexpr int value = 1 + Log2<V / 2>::value; };
template <> struct Log2<1> { static constexpr int value = 0; };

template <int E>
__device__ __forceinline__ void warp_bitonic_sort(uint64_t (&key)[E], int lane) {
  constexpr int LOGN = 5 + Log2<E>::value;
  // counted loops (ls, lt) so that full unrolling -- and with it static register
  // indexing of key[] -- is guaranteed
#pragma unroll
  for (int ls = 1; ls <= LOGN; ++ls) {
#pragma unroll
    for (int lt = ls - 1; lt >= 0; --lt) {
      const int size = 1 << ls, stride = 1 << lt;
      if (lt >= 5) {
        const int es = 1 << (lt - 5);
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if ((e & es) == 0) {
            const bool up = (((e * 32) & size) == 0);  // size >= 64 here: depends on e only
            const uint64_t a = key[e], b = key[e | es];
            const bool sw = (a > b) == up;
            key[e] = sw ? b : a;
            key[e | es] = sw ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const uint64_t other = __shfl_xor_sync(FULL, key[e], stride);
          const int p = e * 32 + lane;
          const bool up = ((p & size) == 0);
          const bool lower = ((lane & stride) == 0);
          const uint64_t mn = key[e] < other ? key[e] : other;
          const uint64_t mx = key[e] < other ? other : key[e];
          key[e] = (lower == up) ? mn : mx;
        }
      }
    }
  }
}

// ---- per-row candidate buffer: sort by (key, idx) and keep the first `keep` -------------
// keys/idxs point at one row's buffer of capacity 32*E holding `cnt` valid entries.
// Called by a full warp with warp-uniform arguments.  Writes the sorted first
// min(cnt, keep) entries back and returns the key of the last kept entry (the new
// admission threshold) -- or +inf if fewer than `keep` entries exist.
template <int E>
__device__ __forceinline__ float warp_sort_truncate(float *keys, uint32_t *idxs, int cnt, int keep,
                                                    int lane, int *new_cnt) {
  uint64_t kk[E];
  __syncwarp();
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int p = e * 32 + lane;
    kk[e] = ~0ull;
    if (p < cnt) kk[e] = (static_cast<uint64_t>(f2ord(keys[p])) << 32) | idxs[p];
  }
  warp_bitonic_sort<E>(kk, lane);
  const int nk = cnt < keep ? cnt : keep;
  __syncwarp();
  float thr = __int_as_float(0x7f800000);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int p = e * 32 + lane;
    if (p < nk) {
      keys[p] = ord2f(static_cast<uint32_t>(kk[e] >> 32));
      idxs[p] = static_cast<uint32_t>(kk[e]);
    }
    // broadcast the key at position keep-1 (if it exists)
    if (cnt >= keep && (keep - 1) / 32 == e) {
      const uint64_t v = __shfl_sync(FULL, kk[e], (keep - 1) & 31);
      thr = ord2f(static_cast<uint32_t>(v >> 32));
    }
  }
  __syncwarp();
  *new_cnt = nk;
  return thr;
}

// ---- per-row candidate buffer: cheap threshold selection (no sort) ------------------------
// Keeps at least the `m` smallest keys of the row's `cnt` entries (at most `keep_max`),
// compacts them to the front PRESERVING their order (so a buffer filled in increasing index
// order stays in index order) and returns the new admission threshold tau with the
// invariant  "every dropped entry has key >= tau".  Requires cnt >= m and m <= keep_max.
//
// The threshold is found by a bitwise binary search on the order-preserving integer image of
// the keys: ~12 warp instructions per bit instead of a ~2.5k-instruction bitonic sort, with
// an early exit as soon as the count of kept entries lands in [m, m + slack].
template <int E>
__device__ __forceinline__ float warp_select_truncate(float *keys, uint32_t *idxs, int cnt, int m, int keep_max,
                                                      int lane, int *new_cnt) {
  uint32_t u[E], id[E];
  __syncwarp();
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int p = e * 32 + lane;
    u[e] = 0xffffffffu;
    id[e] = 0;
    if (p < cnt) {
      u[e] = f2ord(keys[p]);
      id[e] = idxs[p];
    }
  }
  const int slack = (keep_max - m) < 24 ? (keep_max - m) : 24;
  uint32_t T = 0;       // invariant: count(u < T) < m
  uint32_t cut = 0;     // entries with u < cut are kept
  bool exact = true;    // loop ran to the end: T is the m-th smallest key itself
  for (int b = 31; b >= 0; --b) {
    const uint32_t trial = T | (1u << b);
    int c = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c += (u[e] < trial) ? 1 : 0;
    c = __reduce_add_sync(FULL, c);
    if (c < m) {
      T = trial;
    } else if (c <= m + slack) {
      cut = trial;
      exact = false;
      break;
    }
  }
  int ties_allowed = 0;
  if (exact) {
    // count(u < T) < m <= count(u <= T): keep everything below T plus the first ties
    int c_less = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c_less += (u[e] < T) ? 1 : 0;
    c_less = __reduce_add_sync(FULL, c_less);
    cut = T;
    ties_allowed = keep_max - c_less;
  }
  __syncwarp();
  int base = 0, ties_seen = 0;
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int p = e * 32 + lane;
    const bool valid = p < cnt;
    const bool below = valid && (u[e] < cut);
    const bool tie = exact && valid && (u[e] == cut);
    const unsigned tb = __ballot_sync(FULL, tie);
    const bool tie_kept = tie && (ties_seen + __popc(tb & lt_mask) < ties_allowed);
    ties_seen += __popc(tb);
    const bool keep = below || tie_kept;
    const unsigned kb = __ballot_sync(FULL, keep);
    if (keep) {
      const int dst = base + __popc(kb & lt_mask);
      keys[dst] = ord2f(u[e]);
      idxs[dst] = id[e];
    }
    base += __popc(kb);
  }
  __syncwarp();
  *new_cnt = base;
  return ord2f(cut);
}

// ---- the same selection on a list of interleaved (key bits, index) pairs -----------------
// (the tensor-core screening kernel appends candidates with ONE 8-byte store)
// `load(p)` fetches entry p of the input (p < cnt); the kept entries are written to list[0 .. *new_cnt).  The input
// may be `list` itself or span two buffers (the two half lists of a row): every entry is in a register before the
// first one is written.
template <int E, typename Load>
__device__ __forceinline__ float warp_select_truncate_core(Load load, uint2 *list, int cnt, int m, int keep_max, int lane,
                                                           int *new_cnt) {
  uint32_t u[E], id[E];
  __syncwarp();
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int p = e * 32 + lane;
    u[e] = 0xffffffffu;
    id[e] = 0;
    if (p < cnt) {
      const uint2 v = load(p);
      u[e] = f2ord(__uint_as_float(v.x));
      id[e] = v.y;
    }
  }
  // the keys of a row share their leading bits: start the bitwise search below the common prefix
  uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const bool valid = e * 32 + lane < cnt;
    lo = valid ? min(lo, u[e]) : lo;
    hi = valid ? max(hi, u[e]) : hi;
  }
  lo = __reduce_min_sync(FULL, lo);
  hi = __reduce_max_sync(FULL, hi);
  const int top = 31 - __clz(static_cast<int>((lo ^ hi) | 1u));  // highest bit in which two keys differ
  const int slack = (keep_max - m) < 24 ? (keep_max - m) : 24;
  uint32_t T = (top >= 31) ? 0u : (lo & ~((2u << top) - 1u));  // invariant: count(u < T) < m   (count(u < lo) = 0)
  uint32_t cut = 0;     // entries with u < cut are kept
  bool exact = true;    // loop ran to the end: T is the m-th smallest key itself
  for (int b = top; b >= 0; --b) {
    const uint32_t trial = T | (1u << b);
    int c = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c += (u[e] < trial) ? 1 : 0;
    c = __reduce_add_sync(FULL, c);
    if (c < m) {
      T = trial;
    } else if (c <= m + slack) {
      cut = trial;
      exact = false;
      break;
    }
  }
  int ties_allowed = 0;
  if (exact) {
    // count(u < T) < m <= count(u <= T): keep everything below T plus the first ties
    int c_less = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c_less += (u[e] < T) ? 1 : 0;
    c_less = __reduce_add_sync(FULL, c_less);
    cut = T;
    ties_allowed = keep_max - c_less;
  }
  __syncwarp();
  int base = 0, ties_seen = 0;
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int p = e * 32 + lane;
    const bool valid = p < cnt;
    const bool below = valid && (u[e] < cut);
    const bool tie = exact && valid && (u[e] == cut);
    const unsigned tb = __ballot_sync(FULL, tie);
    const bool tie_kept = tie && (ties_seen + __popc(tb & lt_mask) < ties_allowed);
    ties_seen += __popc(tb);
    const bool keep = below || tie_kept;
    const unsigned kb = __ballot_sync(FULL, keep);
    if (keep) {
      const int dst = base + __popc(kb & lt_mask);
      list[dst] = make_uint2(__float_as_uint(ord2f(u[e])), id[e]);
    }
    base += __popc(kb);
  }
  __syncwarp();
  *new_cnt = base;
  return ord2f(cut);
}
// The same selection on entries that are already in registers: entry e of this lane is (key bits kv[e].x, index
// kv[e].y) and takes part iff bit e of `valid` is set; the kept entries are written to list[0 .. *new_cnt) in
// (e, lane) order.  Requires at least m valid entries in the warp.  (halves_merge_kernel loads both half lists of a
// row speculatively, before their lengths have arrived: one DRAM round trip per row instead of two.)
template <int E>
__device__ __forceinline__ float warp_select_truncate_regs(const uint2 (&kv)[E], unsigned valid, uint2 *list, int m,
                                                           int keep_max, int lane, int *new_cnt) {
  uint32_t u[E];
  uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const bool v = (valid >> e) & 1u;
    u[e] = v ? f2ord(__uint_as_float(kv[e].x)) : 0xffffffffu;
    lo = v ? min(lo, u[e]) : lo;
    hi = v ? max(hi, u[e]) : hi;
  }
  lo = __reduce_min_sync(FULL, lo);
  hi = __reduce_max_sync(FULL, hi);
  const int top = 31 - __clz(static_cast<int>((lo ^ hi) | 1u));  // highest bit in which two keys differ
  const int slack = (keep_max - m) < 24 ? (keep_max - m) : 24;
  uint32_t T = (top >= 31) ? 0u : (lo & ~((2u << top) - 1u));  // invariant: count(valid u < T) < m
  uint32_t cut = 0;
  bool exact = true;
  for (int b = top; b >= 0; --b) {
    const uint32_t trial = T | (1u << b);
    int c = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c += (((valid >> e) & 1u) && u[e] < trial) ? 1 : 0;
    c = __reduce_add_sync(FULL, c);
    if (c < m) {
      T = trial;
    } else if (c <= m + slack) {
      cut = trial;
      exact = false;
      break;
    }
  }
  int ties_allowed = 0;
  if (exact) {
    int c_less = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c_less += (((valid >> e) & 1u) && u[e] < T) ? 1 : 0;
    c_less = __reduce_add_sync(FULL, c_less);
    cut = T;
    ties_allowed = keep_max - c_less;
  }
  int base = 0, ties_seen = 0;
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const bool v = (valid >> e) & 1u;
    const bool below = v && (u[e] < cut);
    const bool tie = exact && v && (u[e] == cut);
    const unsigned tb = __ballot_sync(FULL, tie);
    const bool tie_kept = tie && (ties_seen + __popc(tb & lt_mask) < ties_allowed);
    ties_seen += __popc(tb);
    const bool keep = below || tie_kept;
    const unsigned kb = __ballot_sync(FULL, keep);
    if (keep) list[base + __popc(kb & lt_mask)] = kv[e];
    base += __popc(kb);
  }
  __syncwarp();
  *new_cnt = base;
  return ord2f(cut);
}

template <int E>
__device__ __forceinline__ float warp_select_truncate_pairs(uint2 *list, int cnt, int m, int keep_max, int lane,
                                                            int *new_cnt) {
  return warp_select_truncate_core<E>([&](int p) { return list[p]; }, list, cnt, m, keep_max, lane, new_cnt);
}
// the input is the concatenation of list[0 .. cnt_a) and list_b[0 .. cnt_b); the result goes to `list`
template <int E>
__device__ __forceinline__ float warp_select_truncate_pairs_2(uint2 *list, int cnt_a, const uint2 *list_b, int cnt_b, int m,
                                                              int keep_max, int lane, int *new_cnt) {
  return warp_select_truncate_core<E>([&](int p) { return p < cnt_a ? list[p] : list_b[p - cnt_a]; }, list, cnt_a + cnt_b,
                                      m, keep_max, lane, new_cnt);
}

// ---- m-th smallest of a list of floats: threshold only ------------------------------------
// Returns a value tau with  count(keys < tau) >= m  (and at most m + slack such keys when the search
// stops early, exactly the m-th smallest key's successor otherwise).  Requires cnt >= m.  No list is
// rewritten: this serves the threshold pre-pass of the screening kernel, where only the value matters.
template <int E>
__device__ __forceinline__ float warp_select_threshold(const float *keys, int cnt, int m, int slack, int lane) {
  uint32_t u[E];
  __syncwarp();
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int p = e * 32 + lane;
    u[e] = (p < cnt) ? f2ord(keys[p]) : 0xffffffffu;
  }
  uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const bool valid = e * 32 + lane < cnt;
    lo = valid ? min(lo, u[e]) : lo;
    hi = valid ? max(hi, u[e]) : hi;
  }
  lo = __reduce_min_sync(FULL, lo);
  hi = __reduce_max_sync(FULL, hi);
  const int top = 31 - __clz(static_cast<int>((lo ^ hi) | 1u));
  uint32_t T = (top >= 31) ? 0u : (lo & ~((2u << top) - 1u));  // invariant: count(u < T) < m
  for (int b = top; b >= 0; --b) {
    const uint32_t trial = T | (1u << b);
    int c = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c += (u[e] < trial) ? 1 : 0;
    c = __reduce_add_sync(FULL, c);
    if (c < m) T = trial;
    else if (c <= m + slack) return ord2f(trial);
  }
  // T is the m-th smallest key itself: count(u < T) < m <= count(u <= T)
  return (T < 0xff800000u) ? ord2f(T + 1u) : ord2f(T);
}

// ---- the same for several rows in a row: the keys of row r+1 are fetched while row r is searched ----------
// keys of row r: base + r * stride (floats), `cnt` of them (cnt >= m); out[r] = threshold, r in [r0, r1).
template <int E>
__device__ __forceinline__ void warp_select_threshold_rows(const float *base, size_t stride, int r0, int r1, int cnt,
                                                           int m, int slack, int lane, float *out) {
  if (r0 >= r1) return;
  uint32_t u[E], un[E];
  auto fetch = [&](int r, uint32_t(&dst)[E]) {
    const float *keys = base + static_cast<size_t>(r) * stride;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int p = e * 32 + lane;
      dst[e] = (p < cnt) ? f2ord(keys[p]) : 0xffffffffu;
    }
  };
  fetch(r0, un);
  for (int r = r0; r < r1; ++r) {
#pragma unroll
    for (int e = 0; e < E; ++e) u[e] = un[e];
    if (r + 1 < r1) fetch(r + 1, un);
    uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const bool valid = e * 32 + lane < cnt;
      lo = valid ? min(lo, u[e]) : lo;
      hi = valid ? max(hi, u[e]) : hi;
    }
    lo = __reduce_min_sync(FULL, lo);
    hi = __reduce_max_sync(FULL, hi);
    const int top = 31 - __clz(static_cast<int>((lo ^ hi) | 1u));
    uint32_t T = (top >= 31) ? 0u : (lo & ~((2u << top) - 1u));  // invariant: count(u < T) < m
    uint32_t res = 0u;
    bool found = false;
    for (int b = top; b >= 0; --b) {
      const uint32_t trial = T | (1u << b);
      int c = 0;
#pragma unroll
      for (int e = 0; e < E; ++e) c += (u[e] < trial) ? 1 : 0;
      c = __reduce_add_sync(FULL, c);
      if (c < m) {
        T = trial;
      } else if (c <= m + slack) {
        res = trial;
        found = true;
        break;
      }
    }
    if (!found) res = (T < 0xff800000u) ? T + 1u : T;  // T is the m-th smallest key itself
    if (lane == 0) out[r] = ord2f(res);
  }
  __syncwarp();
}

// ---- the same on 15-bit keys, two per register -------------------------------------------------------------------
// The search above costs one compare + two dependent integer steps per key and bit (that is how `c += (u < trial)`
// compiles) over up to 32 bits, and with four warps per scheduler the screening kernel's threshold selection was bound
// by exactly that instruction count.  A STARTING threshold only has to bound the m-th smallest key from above, so the
// keys of a row are quantised linearly to 15 bits over [min, max] of the row (q = floor((v - lo) * 32766 / range),
// monotone), packed two per register with a guard bit each, and one subtraction of the packed trial value compares
// both: the guard bit of a half survives iff q >= trial.  The guard bits of the row's E / 2 registers are gathered
// into one word (shift + mask-or per register) and counted with a single POPC: 1.5 instructions per key and bit, at
// most 15 bits.  The result is rounded UP to the next quantum (range / 32766: a fraction of a key unit).
// keys of row r: base + r * stride (floats), `cnt` of them (cnt >= m, cnt <= 32 * E); out[r] = a value with at least m
// of the row's keys strictly below it.  Keys >= ignore_above (the minima of all-padding half tiles, which would stretch
// the range and coarsen the quantum) are left out; should fewer than m keys remain, the result is just above the
// largest of them -- too tight to be a bound, which costs such a row the refine pass, never exactness.
template <int E>
__device__ __forceinline__ void warp_select_threshold_rows_q15(const float *base, size_t stride, int r0, int r1, int cnt,
                                                               int m, int slack, float ignore_above, int lane,
                                                               float *out) {
  static_assert(E % 2 == 0 && E <= 16, "two keys per register, eight guard-bit pairs per word");
  if (r0 >= r1) return;
  const float inf = __int_as_float(0x7f800000);
  float v[E], vn[E];
  auto fetch = [&](int r, float(&dst)[E]) {
    const float *keys = base + static_cast<size_t>(r) * stride;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int p = e * 32 + lane;
      dst[e] = (p < cnt) ? keys[p] : inf;
    }
  };
  fetch(r0, vn);
  for (int r = r0; r < r1; ++r) {
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = vn[e];
    if (r + 1 < r1) fetch(r + 1, vn);
    float lo = inf, hi = -inf;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      lo = fminf(lo, v[e]);
      hi = (v[e] < ignore_above) ? fmaxf(hi, v[e]) : hi;  // (absent keys are +inf)
    }
    lo = ord2f(__reduce_min_sync(FULL, f2ord(lo)));
    hi = ord2f(__reduce_max_sync(FULL, f2ord(hi)));
    const float range = hi - lo;
    if (!(range > 0.0f) || !(range < inf)) {  // all keys equal (or not finite): anything above the largest will do
      if (lane == 0) out[r] = (hi < inf && hi > -inf) ? hi + fabsf(hi) * 2.4e-7f + 1e-30f : inf;
      continue;
    }
    const float scale = 32766.0f / range;
    uint32_t x[E / 2];
#pragma unroll
    for (int e = 0; e < E; e += 2) {
      const int q0 = (v[e] < ignore_above) ? min(static_cast<int>((v[e] - lo) * scale), 32766) : 32767;
      const int q1 = (v[e + 1] < ignore_above) ? min(static_cast<int>((v[e + 1] - lo) * scale), 32766) : 32767;
      x[e / 2] = (static_cast<uint32_t>(q1 | 0x8000) << 16) | static_cast<uint32_t>(q0 | 0x8000);
    }
    uint32_t T = 0u, res = 0u;  // invariant: count(q < T) < m
    bool found = false;
    for (int b = 14; b >= 0; --b) {
      const uint32_t trial = T | (1u << b);
      const uint32_t tt = trial * 0x00010001u;
      uint32_t bits = 0u;
#pragma unroll
      for (int j = 0; j < E / 2; ++j) bits = (bits >> 1) | ((x[j] - tt) & 0x80008000u);  // guard bit kept: q >= trial
      const int c = __reduce_add_sync(FULL, E - __popc(bits));
      if (c < m) {
        T = trial;
      } else if (c <= m + slack) {
        res = trial;
        found = true;
        break;
      }
    }
    if (!found) res = T + 1u;  // count(q < T + 1) = count(q <= T) >= m
    // q < res  =>  (v - lo) * scale < res (up to the rounding of that product, < 0.01 quantum)  =>  v < lo + res / scale
    if (lane == 0) {
      const float t = lo + (static_cast<float>(res) + 0.02f) * (range * (1.0f / 32766.0f));
      out[r] = t + fabsf(t) * 2.4e-7f;
    }
  }
  __syncwarp();
}
#endif  // __CUDACC__

}  // namespace vvb
