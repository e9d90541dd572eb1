// Running top-k merge for the database-sharded ("ring") graph build.
//
// A rank owns query rows and one shard of the database; the other shards arrive one per ring step.
// Each step runs the ordinary graph builder on [shard ; own rows] (stacked in GLOBAL row order, so
// that the builder's (d2, local index) tie-break agrees with the contract's (d2, global index)) and
// this kernel folds the step's neighbour lists into the rank's running lists:
//
//   * entries of the new list outside the local range [keep_lo, keep_hi) are dropped -- in a ring
//     step those are the rank's own rows, which step 0 has already ranked;
//   * kept entries get their global index (local + global_off) and the contract's distance
//     (left-to-right FP32 fmaf chain, recomputed here: the builder returns sqrtf(d2), which merges
//     neighbouring d2 values and cannot be used to order candidates of different steps);
//   * both lists are ascending in the 64-bit key (d2 bits, global index); each element's rank in the
//     merged order is its own position plus a binary search in the other list, and the first k are
//     written back.  Every member of the global top-k of a row is in the top-k of its own step
//     (fewer than k rows beat it anywhere, hence also among [own ; shard]), so the final lists are
//     bit-identical to a build over the whole matrix.
//
// One warp per query row; lists live in shared memory.
#include "common.cuh"

namespace vvb {

constexpr uint64_t MERGE_INVALID = ~0ull;

__device__ __forceinline__ int lower_bound_u64(const uint64_t *a, int n, uint64_t v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(128) knn_merge_kernel(
    const float *__restrict__ xc, int d, int64_t q0, int64_t nq, const int64_t *__restrict__ new_idx, int kn,
    int64_t keep_lo, int64_t keep_hi, int64_t global_off, int64_t *__restrict__ run_idx,
    float *__restrict__ run_d2, int k) {
  extern __shared__ uint64_t lists[];  // per warp: k running keys, kn new keys, k merged keys
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t t = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + wib;
  if (t >= nq) return;  // warp-uniform
  uint64_t *run = lists + static_cast<size_t>(wib) * (2 * k + kn);
  uint64_t *nw = run + k;
  uint64_t *out = nw + kn;
  int64_t *ri = run_idx + t * (k + 1);
  float *rd = run_d2 + t * (k + 1);
  const float *xq = xc + (q0 + t) * d;
  for (int p = lane; p < k; p += 32) {
    const int64_t j = ri[p + 1];
    run[p] = j < 0 ? MERGE_INVALID : ((static_cast<uint64_t>(__float_as_uint(rd[p + 1])) << 32) | static_cast<uint32_t>(j));
    out[p] = MERGE_INVALID;
  }
  // new entries: the dropped ones leave holes; positions are compacted with a warp prefix sum so that
  // the list stays ascending (the builder emits it ascending in (d2, local index) = (d2, global index))
  int n_new = 0;
  for (int p0 = 0; p0 < kn; p0 += 32) {
    const int p = p0 + lane;
    int64_t j = -1;
    if (p < kn) j = new_idx[t * (kn + 1) + p + 1];
    const bool keep = j >= keep_lo && j < keep_hi;
    uint64_t key = MERGE_INVALID;
    if (keep) {
      const float *y = xc + j * d;
      float acc = 0.0f;
      for (int c = 0; c < d; ++c) {  // the contract's chain: left to right, one rounding per fmaf
        const float df = __fsub_rn(__ldg(xq + c), __ldg(y + c));
        acc = __fmaf_rn(df, df, acc);
      }
      key = (static_cast<uint64_t>(__float_as_uint(acc)) << 32) | static_cast<uint32_t>(j + global_off);
    }
    const unsigned m = __ballot_sync(FULL, keep);
    if (keep) nw[n_new + __popc(m & ((1u << lane) - 1u))] = key;
    n_new += __popc(m);
  }
  __syncwarp();
  int n_run = 0;  // valid running entries form a prefix
  for (int p0 = 0; p0 < k; p0 += 32) {
    const int p = p0 + lane;
    n_run += __popc(__ballot_sync(FULL, p < k && run[p] != MERGE_INVALID));
  }
  // keys are distinct (an index is in at most one shard, and own rows are dropped after step 0)
  for (int p = lane; p < n_run; p += 32) {
    const int r = p + lower_bound_u64(nw, n_new, run[p]);
    if (r < k) out[r] = run[p];
  }
  for (int p = lane; p < n_new; p += 32) {
    const int r = p + lower_bound_u64(run, n_run, nw[p]);
    if (r < k) out[r] = nw[p];
  }
  __syncwarp();
  for (int p = lane; p < k; p += 32) {
    const uint64_t key = out[p];
    if (key == MERGE_INVALID) {
      ri[p + 1] = -1;
      rd[p + 1] = __int_as_float(0x7f800000);
    } else {
      ri[p + 1] = static_cast<int64_t>(static_cast<uint32_t>(key));
      rd[p + 1] = __uint_as_float(static_cast<uint32_t>(key >> 32));
    }
  }
}

__global__ void knn_merge_init_kernel(int64_t *__restrict__ run_idx, float *__restrict__ run_d2, int64_t nq, int k,
                                      int64_t self0) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= nq * (k + 1)) return;
  const int64_t row = e / (k + 1);
  const int col = static_cast<int>(e - row * (k + 1));
  run_idx[e] = col == 0 ? self0 + row : -1;
  run_d2[e] = col == 0 ? 0.0f : __int_as_float(0x7f800000);
}

__global__ void knn_merge_finish_kernel(const float *__restrict__ run_d2, int64_t count, float *__restrict__ dist) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e < count) dist[e] = __fsqrt_rn(run_d2[e]);  // knn.py:135-136 reports distances, not squares
}

int launch_knn_merge(const float *xc, int64_t nc, int d, int64_t q0, int64_t nq, const int64_t *new_idx, int kn,
                     int64_t keep_lo, int64_t keep_hi, int64_t global_off, int64_t *run_idx, float *run_d2, int k,
                     cudaStream_t st) {
  VVB_REQUIRE(nq >= 0 && q0 >= 0 && q0 + nq <= nc && d >= 1, "knn merge: query rows outside the stacked matrix");
  VVB_REQUIRE(k >= 1 && kn >= 0 && kn <= k, "knn merge: need 0 <= kn <= k and k >= 1");
  VVB_REQUIRE(keep_lo >= 0 && keep_hi <= nc && keep_lo <= keep_hi, "knn merge: kept range outside the stacked matrix");
  VVB_REQUIRE(keep_hi + global_off <= 0xffffffffll && keep_lo + global_off >= 0, "knn merge: global indices must fit 32 bits");
  if (nq == 0 || kn == 0 || keep_lo == keep_hi) return VVB_OK;
  const int warps = 4;
  const size_t smem = static_cast<size_t>(warps) * (2 * k + kn) * sizeof(uint64_t);
  if (smem > 200 * 1024) {
    set_error("knn merge: k = %d exceeds the shared-memory list capacity", k);
    return VVB_ENOTSUP;
  }
  if (smem > 48 * 1024)
    VVB_CUDA(cudaFuncSetAttribute(knn_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const unsigned grid = static_cast<unsigned>((nq + warps - 1) / warps);
  knn_merge_kernel<<<grid, warps * 32, smem, st>>>(xc, d, q0, nq, new_idx, kn, keep_lo, keep_hi, global_off, run_idx,
                                                   run_d2, k);
  VVB_CHECK_LAUNCH();
  return VVB_OK;
}

int launch_knn_merge_init(int64_t *run_idx, float *run_d2, int64_t nq, int k, int64_t self0, cudaStream_t st) {
  VVB_REQUIRE(nq >= 0 && k >= 1 && self0 >= 0, "knn merge init: bad sizes");
  const int64_t count = nq * (k + 1);
  if (count == 0) return VVB_OK;
  knn_merge_init_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, st>>>(run_idx, run_d2, nq, k, self0);
  VVB_CHECK_LAUNCH();
  return VVB_OK;
}

int launch_knn_merge_finish(const float *run_d2, int64_t count, float *dist, cudaStream_t st) {
  VVB_REQUIRE(count >= 0, "knn merge finish: bad size");
  if (count == 0) return VVB_OK;
  knn_merge_finish_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, st>>>(run_d2, count, dist);
  VVB_CHECK_LAUNCH();
  return VVB_OK;
}

}  // namespace vvb
