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
#endif  // __CUDACC__

}  // namespace vvb
