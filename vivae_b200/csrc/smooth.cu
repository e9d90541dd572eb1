// kNN smoother -- reproduces /root/reference/src/vivae/knn.py:173-185 on the GPU.
//
// The reference's sweep is in place (res aliases y), so row i reads rows j < i that
// were ALREADY updated in the same sweep (Gauss-Seidel).  That
// dependency DAG only points to lower row indices, which permits a sync-free design:
//   * two arrays per sweep: `prev` (values before the sweep) and `next` (after);
//     row i reads neighbour j from `next` if j < i (after j's done-flag is published)
//     and from `prev` otherwise -- exactly what the sequential loop observes;
//   * one warp per row, rows handed out in increasing order through an atomic ticket,
//     so a warp only ever waits on rows whose owner is already running: no deadlock,
//     whatever the grid size or the row order of the data.
//   * publish / wait protocol (PTX memory model, gpu scope): the owner's 32 lanes store the
//     row, meet at a warp barrier, and ONE lane publishes with st.release.gpu (release is
//     cumulative over the barrier); a waiting lane polls with ld.relaxed.gpu -- no L1
//     invalidation per poll -- and the warp executes ONE fence.acq_rel.gpu once all its
//     polls have succeeded, then a warp barrier, before any lane reads the published rows
//     (with ld.global.cg: L2 is the point of coherence).
//   * tried and dropped: prefetching every neighbour row into L2 before the polling starts
//     (3.3 ms against 2.96 ms at 1M x 50, k=50: the sweep is bound by instruction issue).
// Arithmetic per element matches NumPy's: sequential sum over the k gathered rows in
// neighbour order (starting from the first row, not from zero), true division by k,
// then r + (target - r) * coef as three separately rounded operations (no FMA).
// VVB_SMOOTH_JACOBI (every row reads `prev`) is the documented non-parity variant.
#include <cstdlib>

#include "common.cuh"

namespace vvb {

constexpr int SM_THREADS = 256;
constexpr int SM_ROWS_PER_TICKET_DEFAULT = 1;  // rows a warp claims per atomic ticket (env VVB_SMOOTH_RPT overrides)
constexpr int SM_UNROLL = 16;  // neighbour rows in flight per warp

template <typename T> struct Arith;
template <> struct Arith<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Arith<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

template <typename T, int V> struct VecOf;
template <> struct VecOf<float, 1> { using type = float; };
template <> struct VecOf<double, 1> { using type = double; };
template <> struct VecOf<float, 2> { using type = float2; };
template <> struct VecOf<double, 2> { using type = double2; };
template <typename T> __device__ __forceinline__ T vget(const T &v, int) { return v; }
__device__ __forceinline__ float vget(const float2 &v, int e) { return e ? v.y : v.x; }
__device__ __forceinline__ double vget(const double2 &v, int e) { return e ? v.y : v.x; }

// base + j * ld_bytes as ONE instruction (IMAD.WIDE: 32 x 32 -> 64-bit product plus a 64-bit addend); written as
// (int64) j * stride the compiler hoists the sign-extended stride and emits a four-instruction 64-bit multiply.
template <typename P>
__device__ __forceinline__ const P *row_ptr(const P *base, int j, int ld_bytes) {
  unsigned long long r;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(j), "r"(ld_bytes), "l"(reinterpret_cast<unsigned long long>(base)));
  return reinterpret_cast<const P *>(r);
}

// V   = elements per load (2 when d is even: rows are then 8 / 16-byte aligned and a lane fetches two columns of a
//       neighbour row with ONE load -- the sweep is bound by instruction issue)
// DPL = loads per lane and neighbour row held in registers (a warp covers 32 * DPL * V columns per pass)
// KPL = neighbour indices per lane held in registers (k <= 32 * KPL); indices are narrowed to 32 bits once read
//       (n <= 2^31 - 1 is checked on the host side of this kernel)
template <typename T, int V, int DPL, int KPL, bool GS>
__global__ void __launch_bounds__(SM_THREADS) smooth_sweep_kernel(
    const T *__restrict__ prev, T *next, const int64_t *__restrict__ idx, int64_t ld_idx, int64_t n,
    int d, int64_t ld, int k, T coef, int *flags, int epoch, unsigned long long *ticket, int *bad_index,
    int rows_per_ticket, int64_t row_lo, int64_t row_hi, int jitter) {
  // d columns are processed; rows are `ld` elements apart (ld > d: a column slice of a wider matrix -- the columns of
  // a sweep are independent of each other, which is how several GPUs share one sweep: vvb_smooth_cols_device)
  using TV = typename VecOf<T, V>::type;
  const int lane = threadIdx.x & 31;
  const T kk = static_cast<T>(k);
  const int dv = d / V;  // vectors per row (V == 2 only when d is even)
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = static_cast<unsigned long long>(row_lo) + atomicAdd(ticket, 1ull) * rows_per_ticket;
    base = __shfl_sync(FULL, base, 0);
    if (base >= static_cast<unsigned long long>(row_hi)) break;
    const int64_t row_end = min(static_cast<int64_t>(base) + rows_per_ticket, row_hi);
    for (int64_t i = static_cast<int64_t>(base); i < row_end; ++i) {
      // the row's neighbour list, coalesced, into registers: lane l holds entries l, l+32, ...
      const int64_t *nb = idx + i * ld_idx;
      const int i32 = static_cast<int>(i);
      int mine[KPL];
#pragma unroll
      for (int q = 0; q < KPL; ++q) {
        const int jj = q * 32 + lane;
        int64_t j = (jj < k) ? __ldg(nb + jj) : i;
        if (static_cast<uint64_t>(j) >= static_cast<uint64_t>(n)) {  // never dereference a bad index
          *bad_index = 1;
          j = i;
        }
        mine[q] = static_cast<int>(j);
      }
      if (GS && jitter == 1) {
        // stress mode (VVB_SMOOTH_JITTER=1, tests only): pseudo-random delays of up to ~16 us between claiming a row,
        // polling and publishing shake the order in which rows finish -- the result must not move by a bit
        __nanosleep((static_cast<unsigned>(i) * 2654435761u >> 18) & 0x3fffu);
      }
      if (GS) {
        // wait until every lower-indexed neighbour has been published in this sweep (relaxed polls,
        // exponential back-off), then one acquire fence for the whole warp
        bool polled = false;
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
          if (mine[q] < i32 && jitter != 2) {  // (VVB_SMOOTH_JITTER=2/3: timing probes, results are garbage)
            polled = true;
            unsigned ns = 32;
            while (ld_relaxed(flags + mine[q]) < epoch && jitter != 3) {
              __nanosleep(ns);
              if (ns < 1024) ns <<= 1;
            }
          }
        }
        if (__any_sync(FULL, polled) && jitter != 4) fence_acq_rel();
        __syncwarp();
      }

      for (int c0 = 0; c0 < dv; c0 += 32 * DPL) {
        T acc[DPL][V];
#pragma unroll
        for (int q = 0; q < DPL; ++q)
#pragma unroll
          for (int e = 0; e < V; ++e) acc[q][e] = T(0);
        bool first = true;
        if constexpr (KPL <= 4) {
          // The hot form (k <= 128).  The sweep is bound by its instruction count (ncu: 1,500 warp instructions per row
          // of 50 neighbours, i.e. ~27 per neighbour: a 64-bit multiply, pointer re-loads, two guards with their
          // branches).  Here a neighbour is shuffle + compare + two selects + IMAD.WIDE + load: the lane's column
          // pointers into both arrays are formed once per row (lanes beyond the row are clamped to its last column
          // and drop what they load), the row stride is a 32-bit multiplier, full batches carry no per-neighbour guard
          // and the "first addend" is resolved at compile time.
          const TV *pl[DPL], *nl[DPL];
#pragma unroll
          for (int q = 0; q < DPL; ++q) {
            const int cc = min(c0 + q * 32 + lane, dv - 1);
            pl[q] = reinterpret_cast<const TV *>(prev) + cc;
            nl[q] = reinterpret_cast<const TV *>(static_cast<const T *>(next)) + cc;
          }
          const int ld_bytes = static_cast<int>(ld * static_cast<int64_t>(sizeof(T)));  // (checked on the host side)
#pragma unroll
          for (int jj0 = 0; jj0 < 32 * KPL; jj0 += SM_UNROLL) {
            if (jj0 < k) {  // warp-uniform
              TV v[SM_UNROLL][DPL];
              if (jj0 + SM_UNROLL <= k) {
#pragma unroll
                for (int u = 0; u < SM_UNROLL; ++u) {
                  const int j = __shfl_sync(FULL, mine[(jj0 + u) / 32], (jj0 + u) & 31);
                  const bool from_next = GS && (j < i32);
#pragma unroll
                  for (int q = 0; q < DPL; ++q) {
                    const TV *src = row_ptr((from_next ? nl[q] : pl[q]), j, ld_bytes);
                    v[u][q] = GS ? __ldcg(src) : __ldg(src);
                  }
                }
#pragma unroll
                for (int u = 0; u < SM_UNROLL; ++u)
#pragma unroll
                  for (int q = 0; q < DPL; ++q)
#pragma unroll
                    for (int e = 0; e < V; ++e)
                      acc[q][e] = (jj0 == 0 && u == 0) ? vget(v[u][q], e) : Arith<T>::add(acc[q][e], vget(v[u][q], e));
              } else {
#pragma unroll
                for (int u = 0; u < SM_UNROLL; ++u) {
                  const int j = __shfl_sync(FULL, mine[(jj0 + u) / 32], (jj0 + u) & 31);
                  const bool from_next = GS && (j < i32);
#pragma unroll
                  for (int q = 0; q < DPL; ++q) {
                    const TV *src = row_ptr((from_next ? nl[q] : pl[q]), j, ld_bytes);
                    v[u][q] = TV();
                    if (jj0 + u < k) v[u][q] = GS ? __ldcg(src) : __ldg(src);
                  }
                }
#pragma unroll
                for (int u = 0; u < SM_UNROLL; ++u) {
                  if (jj0 + u < k) {
#pragma unroll
                    for (int q = 0; q < DPL; ++q)
#pragma unroll
                      for (int e = 0; e < V; ++e)
                        acc[q][e] = (jj0 == 0 && u == 0) ? vget(v[u][q], e) : Arith<T>::add(acc[q][e], vget(v[u][q], e));
                  }
                }
              }
            }
          }
          first = false;
        }
#pragma unroll
        for (int kq = 0; kq < (KPL <= 4 ? 0 : KPL); ++kq) {
          if (kq * 32 < k) {  // warp-uniform
            for (int jb = 0; jb < 32 && kq * 32 + jb < k; jb += SM_UNROLL) {
              TV v[SM_UNROLL][DPL];
#pragma unroll
              for (int u = 0; u < SM_UNROLL; ++u) {
                const int j = __shfl_sync(FULL, mine[kq], (jb + u) & 31);
                if (kq * 32 + jb + u < k) {
                  const bool from_next = GS && (j < i32);
                  const TV *src = reinterpret_cast<const TV *>((from_next ? static_cast<const T *>(next) : prev) +
                                                               static_cast<int64_t>(j) * ld);
#pragma unroll
                  for (int q = 0; q < DPL; ++q) {
                    const int c = c0 + q * 32 + lane;
                    TV val = TV();
                    // rows this sweep has rewritten are read from L2 (the point of coherence), never through L1
                    if (c < dv) val = GS ? __ldcg(src + c) : __ldg(src + c);
                    v[u][q] = val;
                  }
                }
              }
#pragma unroll
              for (int u = 0; u < SM_UNROLL; ++u) {
                if (kq * 32 + jb + u < k) {
#pragma unroll
                  for (int q = 0; q < DPL; ++q)
#pragma unroll
                    for (int e = 0; e < V; ++e)
                      acc[q][e] = first ? vget(v[u][q], e) : Arith<T>::add(acc[q][e], vget(v[u][q], e));
                  first = false;
                }
              }
            }
          }
        }
#pragma unroll
        for (int q = 0; q < DPL; ++q) {
          const int c = c0 + q * 32 + lane;
          if (c < dv) {
#pragma unroll
            for (int e = 0; e < V; ++e) {
              const T r = __ldg(prev + i * ld + c * V + e);
              const T target = Arith<T>::div(acc[q][e], kk);
              const T vec = Arith<T>::sub(target, r);
              const T prod = Arith<T>::mul(vec, coef);
              next[i * ld + c * V + e] = Arith<T>::add(r, prod);
            }
          }
        }
      }
      if (GS && jitter == 1) __nanosleep((static_cast<unsigned>(i) * 40503u >> 3) & 0x1fffu);
      if (GS) {
        __syncwarp();  // all 32 lanes' stores to `next` happen before ...
        if (lane == 0) st_release(flags + i, epoch);  // ... the (cumulative) release that publishes the row
      }
    }
  }
}

constexpr int SM_MAX_CHUNKS = 64;  // row-range launches of one sweep (the pipelined host path), one ticket each

struct SmoothWs {
  int *bad;
  void *scratch;
  int *flags;
  unsigned long long *tickets;  // n_iter sweep tickets, then SM_MAX_CHUNKS chunk tickets
};

static SmoothWs carve_smooth_ws(void *workspace, int elem_size, int64_t n, int d, int n_iter) {
  // workspace: [bad-index flag][scratch (n*d) T][flags n int][tickets n_iter + SM_MAX_CHUNKS u64]
  Carver cv(workspace);
  SmoothWs w;
  w.bad = cv.take<int>(1);
  w.scratch = cv.take<char>(static_cast<size_t>(n) * d * elem_size);
  w.flags = cv.take<int>(static_cast<size_t>(n));
  w.tickets = cv.take<unsigned long long>(static_cast<size_t>(n_iter) + SM_MAX_CHUNKS);
  return w;
}

// Zero the bad-index flag, the publish flags and the tickets: once per smooth call, before any sweep.
int smooth_begin(void *workspace, int elem_size, int64_t n, int d, int n_iter, cudaStream_t st) {
  SmoothWs w = carve_smooth_ws(workspace, elem_size, n, d, n_iter);
  VVB_CUDA(cudaMemsetAsync(w.bad, 0, sizeof(int), st));
  VVB_CUDA(cudaMemsetAsync(w.flags, 0, static_cast<size_t>(n) * sizeof(int), st));
  VVB_CUDA(cudaMemsetAsync(w.tickets, 0, (static_cast<size_t>(n_iter) + SM_MAX_CHUNKS) * sizeof(unsigned long long), st));
  return VVB_OK;
}

// One launch of sweep `epoch` (1-based) over rows [row_lo, row_hi) with its own ticket.
template <typename T, int V, int DPL, int KPL>
static int launch_rows(const T *src, T *dst, int64_t n, int d, int64_t ld, const int64_t *idx, int64_t ld_idx, int k, double coef,
                       int mode, const SmoothWs &w, int epoch, unsigned long long *ticket, int64_t row_lo,
                       int64_t row_hi, cudaStream_t st) {
  int blocks_per_sm = 0;
  auto kern_gs = smooth_sweep_kernel<T, V, DPL, KPL, true>;
  auto kern_ja = smooth_sweep_kernel<T, V, DPL, KPL, false>;
  VVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, mode == VVB_SMOOTH_JACOBI ? kern_ja : kern_gs,
                                                         SM_THREADS, 0));
  if (blocks_per_sm < 1) blocks_per_sm = 1;
  int rpt = SM_ROWS_PER_TICKET_DEFAULT;
  if (const char *e = getenv("VVB_SMOOTH_RPT")) rpt = atoi(e) > 0 ? atoi(e) : rpt;
  const int64_t rows = row_hi - row_lo;
  const int64_t warps_needed = (rows + rpt - 1) / rpt;
  int64_t grid = static_cast<int64_t>(sm_count()) * blocks_per_sm;
  const int64_t max_useful = (warps_needed + SM_THREADS / 32 - 1) / (SM_THREADS / 32);
  if (grid > max_useful) grid = max_useful;
  if (mode != VVB_SMOOTH_JACOBI) {
    // Gauss-Seidel: rows in flight wait on each other, so keep the in-flight window a small
    // fraction of n (waiting warps only burn issue slots and L2 polls)
    const int64_t cap = n / (16 * rpt * (SM_THREADS / 32)) + 1;
    if (grid > cap) grid = cap;
  }
  if (grid < 1) grid = 1;
  const char *je = getenv("VVB_SMOOTH_JITTER");
  const int jitter = je ? atoi(je) : 0;
  if (mode == VVB_SMOOTH_JACOBI)
    kern_ja<<<static_cast<unsigned>(grid), SM_THREADS, 0, st>>>(src, dst, idx, ld_idx, n, d, ld, k, static_cast<T>(coef),
                                                               w.flags, epoch, ticket, w.bad, rpt, row_lo, row_hi, 0);
  else
    kern_gs<<<static_cast<unsigned>(grid), SM_THREADS, 0, st>>>(src, dst, idx, ld_idx, n, d, ld, k, static_cast<T>(coef),
                                                               w.flags, epoch, ticket, w.bad, rpt, row_lo, row_hi, jitter);
  VVB_CHECK_LAUNCH();
  return VVB_OK;
}

// chunk < 0: all n_iter sweeps over all rows (ping-pong so that the LAST sweep lands in `out`: sweep s reads
// a_s, writes b_s).  chunk >= 0: rows [row_lo, row_hi) of a ONE-sweep call, x -> out; the caller issues the
// chunks in increasing row order on one stream after smooth_begin().
template <typename T, int V, int DPL, int KPL>
static int launch_sweeps(const T *x, int64_t n, int d, int64_t ld, int col0, const int64_t *idx, int64_t ld_idx, int k, double coef,
                         int n_iter, int mode, T *out, void *workspace, int chunk, int64_t row_lo, int64_t row_hi,
                         cudaStream_t st) {
  SmoothWs w = carve_smooth_ws(workspace, sizeof(T), n, static_cast<int>(ld), n_iter);
  if (chunk >= 0)
    return launch_rows<T, V, DPL, KPL>(x, out, n, d, ld, idx, ld_idx, k, coef, mode, w, 1, w.tickets + n_iter + chunk, row_lo,
                                    row_hi, st);
  const T *src = x;
  for (int it = 0; it < n_iter; ++it) {
    const bool to_out = ((n_iter - 1 - it) % 2) == 0;
    T *dst = to_out ? out : static_cast<T *>(w.scratch) + col0;  // (the scratch matrix has the full row width)
    int rc = launch_rows<T, V, DPL, KPL>(src, dst, n, d, ld, idx, ld_idx, k, coef, mode, w, it + 1, w.tickets + it, 0, n, st);
    if (rc) return rc;
    src = dst;
  }
  return VVB_OK;
}

size_t smooth_workspace_bytes(int elem_size, int64_t n, int d, int n_iter) {
  return 256 + align_up(static_cast<size_t>(n) * d * elem_size, 256) + align_up(static_cast<size_t>(n) * sizeof(int), 256) +
         align_up((static_cast<size_t>(n_iter) + SM_MAX_CHUNKS) * sizeof(unsigned long long), 256) + 256;
}

static int smooth_dispatch(const void *x, int elem_size, int64_t n, int d, int64_t ld, int col0, const int64_t *idx, int64_t ld_idx, int k,
                           double coef, int n_iter, int mode, void *out, void *workspace, int chunk, int64_t row_lo,
                           int64_t row_hi, cudaStream_t st) {
#define VVB_SM_K(T, V, DPL, xp, op)                                                                                   \
  do {                                                                                                                \
    if (k <= 64)                                                                                                      \
      return launch_sweeps<T, V, DPL, 2>(xp, n, d, ld, col0, idx, ld_idx, k, coef, n_iter, mode, op, workspace, chunk, row_lo,  \
                                         row_hi, st);                                                                 \
    if (k <= 128)                                                                                                     \
      return launch_sweeps<T, V, DPL, 4>(xp, n, d, ld, col0, idx, ld_idx, k, coef, n_iter, mode, op, workspace, chunk, row_lo,  \
                                         row_hi, st);                                                                 \
    return launch_sweeps<T, V, DPL, 32>(xp, n, d, ld, col0, idx, ld_idx, k, coef, n_iter, mode, op, workspace, chunk, row_lo,   \
                                        row_hi, st);                                                                  \
  } while (0)
  // the two-column variant, instantiated for the common shapes only (every instantiation costs seconds of ptxas)
#define VVB_SM_K2(T, DPL, xp, op)                                                                                     \
  do {                                                                                                                \
    if (k <= 64)                                                                                                      \
      return launch_sweeps<T, 2, DPL, 2>(xp, n, d, ld, col0, idx, ld_idx, k, coef, n_iter, mode, op, workspace, chunk, row_lo,  \
                                         row_hi, st);                                                                 \
    return launch_sweeps<T, 2, DPL, 4>(xp, n, d, ld, col0, idx, ld_idx, k, coef, n_iter, mode, op, workspace, chunk, row_lo,    \
                                       row_hi, st);                                                                   \
  } while (0)
  // two columns per load when the rows allow it (d even, 8 / 16-byte aligned bases): half the load instructions
  const bool vec2 = (d % 2 == 0) && (ld % 2 == 0) && (col0 % 2 == 0) && d <= 128 && k <= 128 && reinterpret_cast<uintptr_t>(x) % (2 * elem_size) == 0 &&
                    reinterpret_cast<uintptr_t>(out) % (2 * elem_size) == 0 &&
                    reinterpret_cast<uintptr_t>(carve_smooth_ws(workspace, elem_size, n, static_cast<int>(ld), n_iter).scratch) % (2 * elem_size) == 0 &&
                    !(getenv("VVB_SMOOTH_VEC") && atoi(getenv("VVB_SMOOTH_VEC")) == 0);
  if (elem_size == 4) {
    const float *xf = static_cast<const float *>(x);
    float *of = static_cast<float *>(out);
    if (vec2) {
      if (d <= 64) VVB_SM_K2(float, 1, xf, of);
      VVB_SM_K2(float, 2, xf, of);
    }
    if (d <= 32) VVB_SM_K(float, 1, 1, xf, of);
    if (d <= 64) VVB_SM_K(float, 1, 2, xf, of);
    VVB_SM_K(float, 1, 4, xf, of);
  }
  const double *xd = static_cast<const double *>(x);
  double *od = static_cast<double *>(out);
  if (vec2) {
    if (d <= 64) VVB_SM_K2(double, 1, xd, od);
    VVB_SM_K2(double, 2, xd, od);
  }
  if (d <= 32) VVB_SM_K(double, 1, 1, xd, od);
  if (d <= 64) VVB_SM_K(double, 1, 2, xd, od);
  VVB_SM_K(double, 1, 4, xd, od);
#undef VVB_SM_K
#undef VVB_SM_K2
}

static int smooth_check(const void *x, int elem_size, int64_t n, int d, int64_t ld_idx, int k, int n_iter, int mode,
                        const void *out, size_t workspace_bytes) {
  VVB_REQUIRE(elem_size == 4 || elem_size == 8, "smooth: elem_size must be 4 or 8");
  VVB_REQUIRE(n >= 1 && d >= 1 && k >= 1 && ld_idx >= k && n_iter >= 1, "smooth: bad shape arguments");
  VVB_REQUIRE(mode == VVB_SMOOTH_GAUSS_SEIDEL || mode == VVB_SMOOTH_JACOBI, "smooth: unknown mode %d", mode);
  VVB_REQUIRE(out != x, "smooth: out must not alias x");
  VVB_REQUIRE(static_cast<int64_t>(d) * elem_size <= 0x7fffffffLL, "smooth: rows of more than 2 GB are not supported");
  VVB_REQUIRE(workspace_bytes >= smooth_workspace_bytes(elem_size, n, d, n_iter), "smooth: workspace too small");
  VVB_SUPPORTED(k <= 1024, "smooth: k=%d above the supported maximum of 1024", k);
  return VVB_OK;
}

int launch_smooth(const void *x, int elem_size, int64_t n, int d, const int64_t *idx, int64_t ld_idx, int k,
                  double coef, int n_iter, int mode, void *out, void *workspace, size_t workspace_bytes,
                  cudaStream_t st) {
  int rc = smooth_check(x, elem_size, n, d, ld_idx, k, n_iter, mode, out, workspace_bytes);
  if (rc) return rc;
  rc = smooth_begin(workspace, elem_size, n, d, n_iter, st);
  if (rc) return rc;
  return smooth_dispatch(x, elem_size, n, d, d, 0, idx, ld_idx, k, coef, n_iter, mode, out, workspace, -1, 0, n, st);
}

// Rows [row_lo, row_hi) of a one-sweep smooth (chunk = 0, 1, ... < smooth_max_chunks(), increasing row ranges,
// all on one stream, after smooth_begin): lets the host entry point overlap the upload of the neighbour matrix and
// the download of the result with the sweep itself.  In Gauss-Seidel mode a row only reads rows below it from the
// output, so a chunk never needs anything a later chunk produces.
int launch_smooth_rows(const void *x, int elem_size, int64_t n, int d, const int64_t *idx, int64_t ld_idx, int k,
                       double coef, int mode, void *out, void *workspace, size_t workspace_bytes, int chunk,
                       int64_t row_lo, int64_t row_hi, cudaStream_t st) {
  int rc = smooth_check(x, elem_size, n, d, ld_idx, k, 1, mode, out, workspace_bytes);
  if (rc) return rc;
  VVB_REQUIRE(chunk >= 0 && chunk < SM_MAX_CHUNKS && row_lo >= 0 && row_lo <= row_hi && row_hi <= n,
              "smooth: bad row chunk");
  if (row_lo == row_hi) return VVB_OK;
  return smooth_dispatch(x, elem_size, n, d, d, 0, idx, ld_idx, k, coef, 1, mode, out, workspace, chunk, row_lo, row_hi, st);
}

// Columns [col0, col0 + dcols) of all n_iter sweeps over an (n, d) matrix: the sweep treats every column on its own
// (row i's new value in column c depends on column c of its neighbours only), so column slices are independent --
// whole sweeps, Gauss-Seidel order included, and each slice is bit-identical to the same columns of the full call.
// `x` and `out` are the FULL matrices; only the slice of `out` is written.
int launch_smooth_cols(const void *x, int elem_size, int64_t n, int d, int col0, int dcols, const int64_t *idx,
                       int64_t ld_idx, int k, double coef, int n_iter, int mode, void *out, void *workspace,
                       size_t workspace_bytes, cudaStream_t st) {
  int rc = smooth_check(x, elem_size, n, d, ld_idx, k, n_iter, mode, out, workspace_bytes);
  if (rc) return rc;
  VVB_REQUIRE(col0 >= 0 && dcols >= 0 && col0 + dcols <= d, "smooth: bad column slice [%d, %d) of %d", col0, col0 + dcols, d);
  rc = smooth_begin(workspace, elem_size, n, d, n_iter, st);
  if (rc || dcols == 0) return rc;
  const char *xs = static_cast<const char *>(x) + static_cast<size_t>(col0) * elem_size;
  char *os = static_cast<char *>(out) + static_cast<size_t>(col0) * elem_size;
  return smooth_dispatch(xs, elem_size, n, dcols, d, col0, idx, ld_idx, k, coef, n_iter, mode, os, workspace, -1, 0, n, st);
}

int smooth_max_chunks() { return SM_MAX_CHUNKS; }

}  // namespace vvb
