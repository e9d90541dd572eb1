// Exact FP32 brute-force kNN (the exactness fallback of the tensor-core path, and a
// complete make_knn on its own).
//
// Contract (include/vivae_b200.h): neighbours of row i are all
// j != i ranked by (d2, j), d2 = acc after `for c: t = x_ic - x_jc; acc = fmaf(t,t,acc)`
// in FP32, evaluated left to right.  Zero-padding the coordinate loop is exact
// (fmaf(0,0,acc) == acc), which is what lets the kernel work in 16-column chunks.
//
// Layout: one thread owns one query row (its running threshold `tau` and candidate
// count live in registers), a CTA of 128 threads sweeps the database in tiles of
// 32 rows x 16 columns staged in shared memory and read back as warp-wide
// broadcasts.  Candidates that beat `tau` are appended to the row's buffer in
// global memory (L2 resident); when a buffer is nearly full the warp sorts it
// cooperatively in registers (bitonic, 64-bit (key,idx) composites), keeps the best
// k and tightens `tau`.  Because the database is swept in increasing j and admission
// is a strict `<`, ties are broken towards the lower index, as the contract demands.
#include "common.cuh"

namespace vvb {

constexpr int BR_THREADS = 128;
constexpr int BR_TJ = 32;  // database rows per tile
constexpr int BR_DC = 16;  // coordinates per chunk

template <int E>
__global__ void __launch_bounds__(BR_THREADS) knn_brute_kernel(
    const float *__restrict__ x, int64_t n, int d, const int64_t *__restrict__ qrows, int64_t q0,
    int64_t nq_max, const int *__restrict__ nq_dev, int k, float *__restrict__ cand_key,
    uint32_t *__restrict__ cand_idx, int64_t *__restrict__ out_idx, float *__restrict__ out_dist,
    int64_t out_row0) {
  constexpr int CAP = 32 * E;
  __shared__ __align__(16) float ys[BR_TJ][BR_DC];
  // the row count may live on the device (fallback list of the tensor-core path)
  int64_t nq = nq_max;
  if (nq_dev != nullptr) nq = min(nq_max, static_cast<int64_t>(*nq_dev));
  if (static_cast<int64_t>(blockIdx.x) * BR_THREADS >= nq) return;  // whole CTA idle

  const int lane = threadIdx.x & 31;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * BR_THREADS + threadIdx.x;
  const bool active = t < nq;
  const int64_t grow = active ? (qrows ? qrows[t] : q0 + t) : -1;
  const float *xq = x + (active ? grow : 0) * d;
  float *ck = cand_key + t * CAP;  // workspace is sized for gridDim.x * BR_THREADS rows
  uint32_t *ci = cand_idx + t * CAP;

  float tau = active ? __int_as_float(0x7f800000) : -__int_as_float(0x7f800000);
  int cnt = 0;

  for (int64_t j0 = 0; j0 < n; j0 += BR_TJ) {
    // ---- make room for up to BR_TJ admissions (warp-cooperative sort + truncate)
    unsigned need = __ballot_sync(FULL, cnt > CAP - BR_TJ);
    while (need) {
      const int src = __ffs(need) - 1;
      need &= need - 1;
      const int64_t tt = __shfl_sync(FULL, t, src);
      const int c = __shfl_sync(FULL, cnt, src);
      int nc;
      const float thr =
          warp_select_truncate<E>(cand_key + tt * CAP, cand_idx + tt * CAP, c, k, CAP - 2 * BR_TJ, lane, &nc);
      if (lane == src) {
        cnt = nc;
        tau = thr;
      }
    }

    float acc[BR_TJ];
#pragma unroll
    for (int r = 0; r < BR_TJ; ++r) acc[r] = 0.0f;

    for (int c0 = 0; c0 < d; c0 += BR_DC) {
      __syncthreads();  // previous chunk fully consumed
#pragma unroll
      for (int it = 0; it < (BR_TJ * BR_DC) / BR_THREADS; ++it) {
        const int e = it * BR_THREADS + threadIdx.x;
        const int r = e / BR_DC, c = e % BR_DC;
        const int64_t j = j0 + r;
        float v = 0.0f;
        if (j < n && c0 + c < d) v = __ldg(x + j * d + c0 + c);
        ys[r][c] = v;
      }
      float xr[BR_DC];
#pragma unroll
      for (int c = 0; c < BR_DC; ++c) xr[c] = (active && c0 + c < d) ? __ldg(xq + c0 + c) : 0.0f;
      __syncthreads();
#pragma unroll
      for (int r = 0; r < BR_TJ; ++r) {
#pragma unroll
        for (int c4 = 0; c4 < BR_DC / 4; ++c4) {
          const float4 y = *reinterpret_cast<const float4 *>(&ys[r][c4 * 4]);
          float tdiff;
          tdiff = __fsub_rn(xr[c4 * 4 + 0], y.x); acc[r] = __fmaf_rn(tdiff, tdiff, acc[r]);
          tdiff = __fsub_rn(xr[c4 * 4 + 1], y.y); acc[r] = __fmaf_rn(tdiff, tdiff, acc[r]);
          tdiff = __fsub_rn(xr[c4 * 4 + 2], y.z); acc[r] = __fmaf_rn(tdiff, tdiff, acc[r]);
          tdiff = __fsub_rn(xr[c4 * 4 + 3], y.w); acc[r] = __fmaf_rn(tdiff, tdiff, acc[r]);
        }
      }
    }

    // ---- admission: strict `<` against the row's threshold, self excluded
#pragma unroll
    for (int r = 0; r < BR_TJ; ++r) {
      const int64_t j = j0 + r;
      if (acc[r] < tau && j < n && j != grow) {
        ck[cnt] = acc[r];
        ci[cnt] = static_cast<uint32_t>(j);
        ++cnt;
      }
    }
  }

  // ---- finalize: sort every row of this warp and emit [self, k best]
  for (int src = 0; src < 32; ++src) {
    const int64_t tt = __shfl_sync(FULL, t, src);
    const int c = __shfl_sync(FULL, cnt, src);
    const int64_t g = __shfl_sync(FULL, grow, src);
    if (tt >= nq) continue;  // warp-uniform
    float *rk = cand_key + tt * CAP;
    uint32_t *ri = cand_idx + tt * CAP;
    int nc;
    (void)warp_sort_truncate<E>(rk, ri, c, k, lane, &nc);
    int64_t *oi = out_idx + (g - out_row0) * (k + 1);
    float *od = out_dist + (g - out_row0) * (k + 1);
    if (lane == 0) {
      oi[0] = g;
      od[0] = 0.0f;
    }
    for (int p = lane; p < k; p += 32) {
      if (p < nc) {
        oi[p + 1] = ri[p];
        od[p + 1] = __fsqrt_rn(rk[p]);
      } else {  // fewer than k admissible neighbours (NaN input): mark clearly
        oi[p + 1] = -1;
        od[p + 1] = __int_as_float(0x7fc00000);
      }
    }
  }
}

size_t brute_workspace_bytes(int64_t nq, int k) {
  const int64_t rows = (nq + BR_THREADS - 1) / BR_THREADS * BR_THREADS;
  int cap = 128;
  while (cap < k + BR_TJ + 32) cap *= 2;
  return align_up(static_cast<size_t>(rows) * cap * sizeof(float), 256) +
         align_up(static_cast<size_t>(rows) * cap * sizeof(uint32_t), 256) + 256;
}

// Launch the brute-force kernel for `nq` query rows (contiguous from q0, or listed in
// qrows_dev) against the n database rows of x_dev.  Output row of query `g` is g - out_row0.
static int launch_knn_brute_impl(const float *x_dev, int64_t n, int d, int k, const int64_t *qrows_dev,
                                 const int *count_dev, int64_t q0, int64_t nq, int64_t *out_idx, float *out_dist,
                                 int64_t out_row0, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (nq == 0) return VVB_OK;
  VVB_SUPPORTED(n <= 0xffffffffLL, "brute kNN: n=%lld exceeds the 32-bit candidate index", (long long)n);
  VVB_SUPPORTED(k + BR_TJ + 32 <= 2048, "brute kNN: k=%d above the supported maximum of 1984", k);
  VVB_REQUIRE(workspace_bytes >= brute_workspace_bytes(nq, k), "brute kNN: workspace too small");
  const int64_t rows = (nq + BR_THREADS - 1) / BR_THREADS * BR_THREADS;
  int cap = 128;
  while (cap < k + BR_TJ + 32) cap *= 2;
  Carver cv(workspace);
  float *ck = cv.take<float>(static_cast<size_t>(rows) * cap);
  uint32_t *ci = cv.take<uint32_t>(static_cast<size_t>(rows) * cap);
  const unsigned grid = static_cast<unsigned>(rows / BR_THREADS);
#define VVB_BRUTE_CASE(EE)                                                                         \
  case 32 * EE:                                                                                    \
    knn_brute_kernel<EE><<<grid, BR_THREADS, 0, stream>>>(x_dev, n, d, qrows_dev, q0, nq, count_dev, k, ck, \
                                                          ci, out_idx, out_dist, out_row0);        \
    break;
  switch (cap) {
    VVB_BRUTE_CASE(4)
    VVB_BRUTE_CASE(8)
    VVB_BRUTE_CASE(16)
    VVB_BRUTE_CASE(32)
    VVB_BRUTE_CASE(64)
    default:
      set_error("brute kNN: unsupported candidate capacity %d", cap);
      return VVB_ENOTSUP;
  }
#undef VVB_BRUTE_CASE
  VVB_CHECK_LAUNCH();
  return VVB_OK;
}

// ------------------------------------------------------------------------------------------------
// Few query rows (the tensor path's certificate rejects a handful of rows per million): the
// thread-per-query kernel above would leave one CTA sweeping the whole database.  Here a CTA works
// on ONE query row and one slice of the database: every thread owns a database row of the current
// 128-row tile (staged through shared memory, bit-exact chain as above), survivors are appended to the
// (row, slice) list in increasing index order (block-wide prefix, no atomics), and a merge kernel
// sorts the slices' top-k lists of a row together.
// ------------------------------------------------------------------------------------------------
constexpr int FEW_THREADS = 128;
constexpr int FEW_E = 16;              // list capacity 512 per (row, slice)
constexpr int FEW_CAP = 32 * FEW_E;
constexpr int FEW_MAX_ROWS = 2048;     // above this the thread-per-query kernel is the efficient one

__global__ void __launch_bounds__(FEW_THREADS) knn_brute_few_kernel(
    const float *__restrict__ x, int64_t n, int d, const int64_t *__restrict__ qrows, int k, int slices,
    int64_t slice_rows, float *__restrict__ list_key, uint32_t *__restrict__ list_idx, int *__restrict__ list_cnt) {
  __shared__ float tile[FEW_THREADS][BR_DC + 1];
  __shared__ float xq[BR_DC];
  __shared__ int warp_hits[FEW_THREADS / 32];
  __shared__ int s_cnt;
  __shared__ float s_tau;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t g = qrows[blockIdx.x];
  const int64_t list = static_cast<int64_t>(blockIdx.x) * slices + blockIdx.y;
  float *lk = list_key + list * FEW_CAP;
  uint32_t *li = list_idx + list * FEW_CAP;
  const int64_t j_begin = static_cast<int64_t>(blockIdx.y) * slice_rows;
  const int64_t j_end = min(n, j_begin + slice_rows);
  if (tid == 0) {
    s_cnt = 0;
    s_tau = __int_as_float(0x7f800000);
  }
  __syncthreads();
  for (int64_t j0 = j_begin; j0 < j_end; j0 += FEW_THREADS) {
    const int64_t j = j0 + tid;
    float acc = 0.0f;
    for (int c0 = 0; c0 < d; c0 += BR_DC) {
      __syncthreads();
      if (tid < BR_DC) xq[tid] = (c0 + tid < d) ? __ldg(x + g * d + c0 + tid) : 0.0f;
#pragma unroll
      for (int it = 0; it < BR_DC; ++it) {
        const int e = it * FEW_THREADS + tid;
        const int r = e / BR_DC, c = e % BR_DC;
        const int64_t jr = j0 + r;
        tile[r][c] = (jr < j_end && c0 + c < d) ? __ldg(x + jr * d + c0 + c) : 0.0f;
      }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < BR_DC; ++c) {  // zero padding is exact: fmaf(0, 0, acc) == acc
        const float df = __fsub_rn(xq[c], tile[tid][c]);
        acc = __fmaf_rn(df, df, acc);
      }
    }
    // append survivors in increasing index order
    const bool hit = j < j_end && j != g && acc < s_tau;
    const unsigned hb = __ballot_sync(FULL, hit);
    if (lane == 0) warp_hits[warp] = __popc(hb);
    __syncthreads();
    int base = s_cnt, total = 0;
#pragma unroll
    for (int w = 0; w < FEW_THREADS / 32; ++w) {
      if (w < warp) base += warp_hits[w];
      total += warp_hits[w];
    }
    if (hit) {
      const int pos = base + __popc(hb & ((1u << lane) - 1u));
      lk[pos] = acc;
      li[pos] = static_cast<uint32_t>(j);
    }
    __syncthreads();
    if (warp == 0) {
      int cnt = s_cnt + total;
      float tau = s_tau;
      if (cnt > FEW_CAP - FEW_THREADS) {  // room for a full tile of survivors next time
        int nc;
        tau = warp_select_truncate<FEW_E>(lk, li, cnt, k, FEW_CAP - 2 * FEW_THREADS, lane, &nc);
        cnt = nc;
      }
      if (lane == 0) {
        s_cnt = cnt;
        s_tau = tau;
      }
    }
    __syncthreads();
  }
  if (warp == 0) {
    int nc;
    (void)warp_sort_truncate<FEW_E>(lk, li, s_cnt, k, lane, &nc);  // sorted by (d2, index)
    if (lane == 0) list_cnt[list] = nc;
  }
}

// one warp per row: merge the slices' sorted top-k lists (<= 2048 entries) and emit [self, k best]
__global__ void __launch_bounds__(32) knn_merge_few_kernel(const int64_t *__restrict__ qrows, int k, int slices,
                                                          const float *__restrict__ list_key,
                                                          const uint32_t *__restrict__ list_idx,
                                                          const int *__restrict__ list_cnt,
                                                          int64_t *__restrict__ out_idx, float *__restrict__ out_dist,
                                                          int64_t out_row0) {
  constexpr int E = 64;
  const int lane = threadIdx.x;
  const int64_t g = qrows[blockIdx.x];
  uint64_t key[E];
#pragma unroll
  for (int e = 0; e < E; ++e) key[e] = ~0ull;
  // entry p of the concatenation lives at (slice = p / k, position = p % k)
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int p = e * 32 + lane;
    const int sl = p / k, pos = p % k;
    if (sl < slices) {
      const int64_t list = static_cast<int64_t>(blockIdx.x) * slices + sl;
      if (pos < list_cnt[list])
        key[e] = (static_cast<uint64_t>(f2ord(list_key[list * FEW_CAP + pos])) << 32) | list_idx[list * FEW_CAP + pos];
    }
  }
  warp_bitonic_sort<E>(key, lane);
  int64_t *oi = out_idx + (g - out_row0) * (k + 1);
  float *od = out_dist + (g - out_row0) * (k + 1);
  if (lane == 0) {
    oi[0] = g;
    od[0] = 0.0f;
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int p = e * 32 + lane;
    if (p < k) {
      if (key[e] != ~0ull) {
        oi[p + 1] = static_cast<int64_t>(static_cast<uint32_t>(key[e]));
        od[p + 1] = __fsqrt_rn(ord2f(static_cast<uint32_t>(key[e] >> 32)));
      } else {
        oi[p + 1] = -1;
        od[p + 1] = __int_as_float(0x7fc00000);
      }
    }
  }
}

int knn_brute_few_max_rows() { return FEW_MAX_ROWS; }

size_t knn_brute_few_workspace_bytes(int k) {
  const size_t lists = static_cast<size_t>(FEW_MAX_ROWS) + 2 * 148 * 4;
  (void)k;
  return lists * FEW_CAP * 8 + lists * sizeof(int) + 1024;
}

// `count` rows (known on the host, 1 <= count <= FEW_MAX_ROWS) listed in qrows_dev
int launch_knn_brute_few(const float *x_dev, int64_t n, int d, int k, const int64_t *qrows_dev, int count,
                         int64_t *out_idx, float *out_dist, int64_t out_row0, void *workspace, size_t workspace_bytes,
                         cudaStream_t stream) {
  if (count <= 0) return VVB_OK;
  VVB_REQUIRE(count <= FEW_MAX_ROWS && k <= FEW_CAP - 2 * FEW_THREADS, "few-row brute kNN: count=%d k=%d unsupported",
              count, k);
  // slices: enough CTAs to fill the GPU about twice, merged lists must fit 2048 entries per row
  int slices = (2 * sm_count() + count - 1) / count;
  const int max_by_merge = 2048 / k;
  if (slices > max_by_merge) slices = max_by_merge;
  if (slices > 64) slices = 64;
  if (slices < 1) slices = 1;
  int64_t slice_rows = (n + slices - 1) / slices;
  slice_rows = (slice_rows + FEW_THREADS - 1) / FEW_THREADS * FEW_THREADS;
  const size_t lists = static_cast<size_t>(count) * slices;
  VVB_REQUIRE(workspace_bytes >= lists * FEW_CAP * 8 + lists * sizeof(int) + 1024, "few-row brute kNN: workspace too small");
  Carver cv(workspace);
  float *lk = cv.take<float>(lists * FEW_CAP);
  uint32_t *li = cv.take<uint32_t>(lists * FEW_CAP);
  int *lc = cv.take<int>(lists);
  knn_brute_few_kernel<<<dim3(count, slices), FEW_THREADS, 0, stream>>>(x_dev, n, d, qrows_dev, k, slices, slice_rows,
                                                                        lk, li, lc);
  VVB_CHECK_LAUNCH();
  knn_merge_few_kernel<<<count, 32, 0, stream>>>(qrows_dev, k, slices, lk, li, lc, out_idx, out_dist, out_row0);
  VVB_CHECK_LAUNCH();
  return VVB_OK;
}

int launch_knn_brute(const float *x_dev, int64_t n, int d, int k, const int64_t *qrows_dev, int64_t q0,
                     int64_t nq, int64_t *out_idx, float *out_dist, int64_t out_row0, void *workspace,
                     size_t workspace_bytes, cudaStream_t stream) {
  return launch_knn_brute_impl(x_dev, n, d, k, qrows_dev, nullptr, q0, nq, out_idx, out_dist, out_row0, workspace,
                               workspace_bytes, stream);
}

// rows listed in qrows_dev[0 .. *count_dev) (count known only on the device; at most max_rows)
int launch_knn_brute_list(const float *x_dev, int64_t n, int d, int k, const int64_t *qrows_dev,
                          const int *count_dev, int64_t max_rows, int64_t *out_idx, float *out_dist,
                          int64_t out_row0, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  return launch_knn_brute_impl(x_dev, n, d, k, qrows_dev, count_dev, 0, max_rows, out_idx, out_dist, out_row0,
                               workspace, workspace_bytes, stream);
}

}  // namespace vvb
