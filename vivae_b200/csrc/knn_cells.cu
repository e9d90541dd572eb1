// Spatial cells for the exact kNN screening sweep.
//
// The screening kernel (knn_tensor.cu) is bound by the tensor pipe: every (query, database row)
// pair it touches costs 192 fp16 MACs at d = 50.  The only way below that floor is to touch
// fewer pairs -- exactly.  The rows are therefore partitioned into P cells (nearest of P pivot
// rows), sorted by cell, and for every ordered pair of cells (a, b) a LOWER BOUND of the
// distance between any row of a and any row of b is computed:
//
//   * discriminant bound: u_ab = normalise((c_b - c_a) / var) with `var` the pooled within-cell
//     variance per coordinate (the data is PCA-like: a diagonal metric is the natural one).
//     For any unit vector u, |x - y| >= u.(y - x), hence
//         |x - y| >= u_ab.(c_b - c_a) - max_{x in a} u_ab.(x - c_a) - max_{y in b} u_ba.(y - c_b)
//     Elongated clusters that the ball bound cannot separate are separated by this one.
//   * ball bound: |c_b - c_a| - r_a - r_b.
//
// All geometry uses min(d, 64) coordinates -- for d > 64 the 64 columns of largest variance (for PCA
// input these are the leading components, for gene-expression input the most variable genes); a
// projection never increases a distance, so the bounds stay valid for the full rows -- and is carried out on centred FP32 values with an
// explicit safety margin.  A query block sweeps the cells in increasing order of its bound and
// stops at the first cell whose bound exceeds every row's current k-th candidate distance; the
// certificate of the re-rank kernel covers the skipped rows (knn_tensor.cu).  Exactness never
// depends on the quality of the partition: bad cells only mean a longer sweep.
#include "knn_cells.cuh"

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace vvb {

// Packed FP32 pairs (sm_100: FFMA2 executes two IEEE fused multiply-adds per instruction).  The two row x P x 64
// kernels of the preparation (pivot assignment, cell extents) are bound by their FMA issue slots; every lane of a
// packed FMA rounds exactly like a scalar fmaf, so the four accumulation chains below are bit-identical to the
// scalar code they replace.
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void fma_f32x2(unsigned long long &acc, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}


int cells_choose_count(int64_t n) {
  if (const char *e = getenv("VVB_TC_CELLS")) {
    const int v = atoi(e);
    if (v >= 1) return v > CELL_MAX ? CELL_MAX : v;
  }
  if (n < 32768) return 1;
  int p = 1;
  while (static_cast<int64_t>(p) * 4096 < n && p < CELL_MAX) p *= 2;
  return p;
}

int64_t cells_perm_capacity(int64_t n, int P) {
  return (n + CELL_GROUP - 1) / CELL_GROUP * CELL_GROUP + static_cast<int64_t>(P) * CELL_GROUP;
}

size_t cells_carve(CellPlan *plan, void *ws, int64_t n, int d, int P) {
  CellPlan L;
  L.P = P;
  L.dsub = d < CELL_DIMS ? d : CELL_DIMS;
  L.n = n;
  L.n_perm_cap = cells_perm_capacity(n, P);
  Carver cv(ws);
  const size_t PP = static_cast<size_t>(P) * P;
  L.cell_of = cv.take<int>(static_cast<size_t>(n));
  L.cnt = cv.take<int>(P);
  L.off = cv.take<int>(P + 1);
  L.cursor = cv.take<int>(P);
  L.perm = cv.take<int>(static_cast<size_t>(L.n_perm_cap));
  L.sums = cv.take<double>(static_cast<size_t>(P) * CELL_DIMS);
  L.sqs = cv.take<double>(static_cast<size_t>(P) * CELL_DIMS);
  L.cen = cv.take<float>(static_cast<size_t>(P) * CELL_DIMS);
  L.var = cv.take<float>(CELL_DIMS);
  L.cw = cv.take<float>(static_cast<size_t>(P) * CELL_DIMS);
  L.inv_norm = cv.take<float>(PP);
  L.gap = cv.take<float>(PP);
  L.cdist = cv.take<float>(PP);
  L.ext = cv.take<unsigned int>(PP);
  L.rad2 = cv.take<unsigned int>(P);
  L.lb = cv.take<float>(PP);
  L.qcnt = cv.take<int>(P);
  L.qoff = cv.take<int>(P + 1);
  L.qcursor = cv.take<int>(P);
  if (plan) *plan = L;
  return cv.used();
}

// ------------------------------------------------------------------ coordinates used by the geometry
// per-column sum of squared centred values, partial sums in double (same tiling as the column means)
__global__ void __launch_bounds__(256) colvar_partial_kernel(const float *__restrict__ x, int64_t n, int d,
                                                            const float *__restrict__ mu, double *__restrict__ partial) {
  const int cl = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + cl;
  double acc = 0.0;
  if (c < d) {
    const float m = __ldg(mu + c);
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * 4 + sub; r < n; r += static_cast<int64_t>(gridDim.x) * 4) {
      const double v = static_cast<double>(__ldg(x + r * d + c) - m);
      acc += v * v;
    }
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  if (sub == 0)
    partial[(static_cast<size_t>(blockIdx.x) * gridDim.y + blockIdx.y) * 64 + cl] =
        sh[cl] + sh[64 + cl] + sh[128 + cl] + sh[192 + cl];
}
// one CTA: total per column, then the CELL_DIMS columns of largest variance (ties -> lowest column), ascending
__global__ void __launch_bounds__(1024) select_dims_kernel(const double *__restrict__ partial, int nblocks, int nyb,
                                                          int d, int *__restrict__ dims) {
  __shared__ float var[4096];
  __shared__ unsigned long long best[32];
  __shared__ int chosen[CELL_DIMS];
  for (int c = threadIdx.x; c < 4096; c += blockDim.x) {
    double acc = -1.0;
    if (c < d) {
      acc = 0.0;
      for (int b = 0; b < nblocks; ++b) acc += partial[(static_cast<size_t>(b) * nyb + c / 64) * 64 + (c & 63)];
    }
    var[c] = static_cast<float>(acc);
  }
  __syncthreads();
  for (int it = 0; it < CELL_DIMS; ++it) {
    unsigned long long mine = 0ull;
    for (int c = threadIdx.x; c < 4096; c += blockDim.x)
      if (var[c] >= 0.0f) {
        const unsigned long long k = (static_cast<unsigned long long>(__float_as_uint(var[c])) << 32) |
                                     static_cast<unsigned>(4095 - c);
        mine = k > mine ? k : mine;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(FULL, mine, o);
      mine = other > mine ? other : mine;
    }
    if ((threadIdx.x & 31) == 0) best[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long m = 0ull;
      for (int w = 0; w < 32; ++w) m = best[w] > m ? best[w] : m;
      const int c = 4095 - static_cast<int>(static_cast<unsigned>(m));
      chosen[it] = c;
      var[c] = -1.0f;
    }
    __syncthreads();
  }
  // ascending column order (keeps the gathers of a row as local as they can be)
  if (threadIdx.x < CELL_DIMS) {
    const int mine = chosen[threadIdx.x];
    int rank = 0;
    for (int j = 0; j < CELL_DIMS; ++j) rank += (chosen[j] < mine) ? 1 : 0;
    dims[rank] = mine;
  }
}
__global__ void identity_dims_kernel(int *__restrict__ dims) { dims[threadIdx.x] = threadIdx.x; }

// dims[0 .. CELL_DIMS): the coordinates the cell geometry uses.  `partial` needs gridDim.x * ceil(d/64) * 64 doubles.
int cells_select_dims(const float *x, int64_t n, int d, const float *mu, double *partial, int nblocks, int *dims,
                      cudaStream_t st) {
  if (d <= CELL_DIMS) {
    identity_dims_kernel<<<1, CELL_DIMS, 0, st>>>(dims);
    VVB_CHECK_LAUNCH();
    return VVB_OK;
  }
  const int nyb = (d + 63) / 64;
  colvar_partial_kernel<<<dim3(nblocks, nyb), 256, 0, st>>>(x, n, d, mu, partial);
  VVB_CHECK_LAUNCH();
  select_dims_kernel<<<1, 1024, 0, st>>>(partial, nblocks, nyb, d, dims);
  VVB_CHECK_LAUNCH();
  return VVB_OK;
}

// ------------------------------------------------------------------ partition
// Pivot p is row floor((p + 0.5) n / P); its first dsub centred coordinates, gathered once (every CTA of the
// assignment kernel used to redo this gather -- three dependent loads and a double-precision division per
// element -- for all P pivots: a quarter of that kernel's instructions).
__global__ void __launch_bounds__(CELL_DIMS) cell_pivots_kernel(const float *__restrict__ x, int64_t n, int d, int dsub,
                                                               const int *__restrict__ dims,
                                                               const float *__restrict__ mu, int P,
                                                               float *__restrict__ piv) {
  const int p = blockIdx.x, c = threadIdx.x;
  const int64_t pr = static_cast<int64_t>((static_cast<double>(p) + 0.5) * static_cast<double>(n) / P);
  piv[p * CELL_DIMS + c] = (c < dsub) ? __ldg(x + pr * d + __ldg(dims + c)) - __ldg(mu + __ldg(dims + c)) : 0.0f;
}

// Two rows per thread: nearest pivot in the first dsub centred coordinates (ties -> lowest pivot).
// Rows [row0, n) of this launch (n = end of the range: a rank of a multi-GPU build assigns only its row shard).
// The pivot block is read from shared memory as 16-byte broadcasts; a warp-wide 16-byte read occupies the shared-
// memory pipe for four cycles whether or not the lanes share an address, and with ONE row per thread that pipe (one
// per SM, against four FMA pipes) bounded the kernel at a third of the FP32 rate -- so every value read feeds two
// rows, and the multiply-adds are issued in packed pairs.  Each row's four accumulation chains are unchanged.
constexpr int ASSIGN_ROWS = 256;  // rows per CTA of 128 threads
__global__ void __launch_bounds__(128) cell_assign_kernel(const float *__restrict__ x, int64_t row0, int64_t n, int d,
                                                         int dsub, const int *__restrict__ dims,
                                                         const float *__restrict__ mu, int P,
                                                         const float *__restrict__ pivots,
                                                         int *__restrict__ cell_of, int *__restrict__ cnt) {
  __shared__ __align__(16) float piv[128][CELL_DIMS];
  __shared__ float pn[128];
  __shared__ int hist[CELL_MAX];
  const int tid = threadIdx.x;
  for (int i = tid; i < P; i += 128) hist[i] = 0;
  const int64_t row_a = row0 + static_cast<int64_t>(blockIdx.x) * ASSIGN_ROWS + tid, row_b = row_a + 128;
  unsigned long long xa[CELL_DIMS / 2], xb[CELL_DIMS / 2];
#pragma unroll
  for (int c = 0; c < CELL_DIMS; c += 2) {
    float va[2], vb[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool on = c + e < dsub;
      const int col = on ? __ldg(dims + c + e) : 0;
      const float m = on ? __ldg(mu + col) : 0.0f;
      va[e] = (on && row_a < n) ? __ldg(x + row_a * d + col) - m : 0.0f;
      vb[e] = (on && row_b < n) ? __ldg(x + row_b * d + col) - m : 0.0f;
    }
    xa[c / 2] = pack_f32x2(va[0], va[1]);
    xb[c / 2] = pack_f32x2(vb[0], vb[1]);
  }
  float best_a = __int_as_float(0x7f800000), best_b = best_a;
  int best_pa = 0, best_pb = 0;
  for (int p0 = 0; p0 < P; p0 += 128) {
    __syncthreads();
    const int np = min(128, P - p0);
    {
      const float4 *src = reinterpret_cast<const float4 *>(pivots + static_cast<size_t>(p0) * CELL_DIMS);
      float4 *dst = reinterpret_cast<float4 *>(&piv[0][0]);
      for (int i = tid; i < np * (CELL_DIMS / 4); i += 128) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    if (tid < np) {
      float s = 0.0f;
      for (int c = 0; c < CELL_DIMS; ++c) s = fmaf(piv[tid][c], piv[tid][c], s);
      pn[tid] = s;
    }
    __syncthreads();
    for (int pl = 0; pl < np; ++pl) {
      // per row four independent chains (columns 0, 1, 2, 3 mod 4), two per packed accumulator
      unsigned long long a01 = 0ull, a23 = 0ull, b01 = 0ull, b23 = 0ull;
      const ulonglong2 *pv = reinterpret_cast<const ulonglong2 *>(piv[pl]);
#pragma unroll
      for (int c4 = 0; c4 < CELL_DIMS / 4; ++c4) {
        const ulonglong2 q = pv[c4];
        fma_f32x2(a01, xa[2 * c4 + 0], q.x);
        fma_f32x2(a23, xa[2 * c4 + 1], q.y);
        fma_f32x2(b01, xb[2 * c4 + 0], q.x);
        fma_f32x2(b23, xb[2 * c4 + 1], q.y);
      }
      float d0, d1, d2, d3;
      unpack_f32x2(a01, d0, d1);
      unpack_f32x2(a23, d2, d3);
      const float score_a = fmaf(-2.0f, (d0 + d1) + (d2 + d3), pn[pl]);
      unpack_f32x2(b01, d0, d1);
      unpack_f32x2(b23, d2, d3);
      const float score_b = fmaf(-2.0f, (d0 + d1) + (d2 + d3), pn[pl]);
      if (score_a < best_a) {
        best_a = score_a;
        best_pa = p0 + pl;
      }
      if (score_b < best_b) {
        best_b = score_b;
        best_pb = p0 + pl;
      }
    }
  }
  if (row_a < n) {
    cell_of[row_a] = best_pa;
    atomicAdd(&hist[best_pa], 1);
  }
  if (row_b < n) {
    cell_of[row_b] = best_pb;
    atomicAdd(&hist[best_pb], 1);
  }
  __syncthreads();
  for (int i = tid; i < P; i += 128)
    if (hist[i]) atomicAdd(cnt + i, hist[i]);
}

// exclusive scan of the cell sizes (rounded up to `granule` rows) by one CTA of 1024 threads
__global__ void __launch_bounds__(1024) cell_scan_kernel(const int *__restrict__ cnt, int P, int granule,
                                                        int *__restrict__ off, int *__restrict__ cursor) {
  __shared__ int sh[CELL_MAX];
  const int tid = threadIdx.x;
  const int mine = (tid < P) ? (cnt[tid] + granule - 1) / granule * granule : 0;
  sh[tid] = mine;
  __syncthreads();
  for (int o = 1; o < CELL_MAX; o <<= 1) {
    const int v = (tid >= o) ? sh[tid - o] : 0;
    __syncthreads();
    sh[tid] += v;
    __syncthreads();
  }
  if (tid < P) {
    off[tid] = sh[tid] - mine;
    cursor[tid] = sh[tid] - mine;
  }
  if (tid == P - 1) off[P] = sh[tid];
}

// rows: row0 + i, or row_list[i] when a list of (original) rows is given
__global__ void __launch_bounds__(256) cell_scatter_kernel(const int *__restrict__ cell_of, int64_t row0, int64_t rows,
                                                          const int64_t *__restrict__ row_list,
                                                          int *__restrict__ cursor, int *__restrict__ list,
                                                          int64_t list_len) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < rows) {
    const int64_t r = row_list ? row_list[i] : row0 + i;
    const int pos = atomicAdd(cursor + cell_of[r], 1);
    list[pos] = static_cast<int>(r);
  } else if (i < list_len) {
    list[i] = -1;  // tail of a query list (database lists are pre-filled with -1 by a memset: list_len = -1)
  }
}

__global__ void __launch_bounds__(256) cell_hist_kernel(const int *__restrict__ cell_of, int64_t row0, int64_t rows,
                                                       const int64_t *__restrict__ row_list, int P,
                                                       int *__restrict__ cnt) {
  __shared__ int hist[CELL_MAX];
  for (int i = threadIdx.x; i < P; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    atomicAdd(&hist[cell_of[row_list ? row_list[i] : row0 + i]], 1);
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x)
    if (hist[i]) atomicAdd(cnt + i, hist[i]);
}

// ------------------------------------------------------------------ cell geometry
// one CTA per group of 128 permuted rows (a group never straddles cells): coordinate sums and sums of squares
__global__ void __launch_bounds__(256) cell_moments_kernel(const float *__restrict__ x, int d, int dsub,
                                                          const int *__restrict__ dims,
                                                          const float *__restrict__ mu, const int *__restrict__ perm,
                                                          const int *__restrict__ off, int P,
                                                          const int *__restrict__ cell_of, double *__restrict__ sums,
                                                          double *__restrict__ sqs, int part, int parts) {
  __shared__ double sh_s[4][CELL_DIMS], sh_q[4][CELL_DIMS];
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * CELL_GROUP;
  if (p0 >= off[P]) return;
  const int a = cell_of[perm[p0]];
  if (a % parts != part) return;  // another rank owns this cell (multi-GPU build); the sums are all-reduced
  const int c = threadIdx.x & 63, rl = threadIdx.x >> 6;
  double s = 0.0, q = 0.0;
  if (c < dsub) {
    const int col = __ldg(dims + c);
    const float m = __ldg(mu + col);
    for (int i = rl; i < CELL_GROUP; i += 4) {
      const int r = perm[p0 + i];
      if (r >= 0) {
        const double v = static_cast<double>(__ldg(x + static_cast<int64_t>(r) * d + col) - m);
        s += v;
        q += v * v;
      }
    }
  }
  sh_s[rl][c] = s;
  sh_q[rl][c] = q;
  __syncthreads();
  if (rl == 0 && c < dsub) {
    atomicAdd(sums + a * CELL_DIMS + c, sh_s[0][c] + sh_s[1][c] + sh_s[2][c] + sh_s[3][c]);
    atomicAdd(sqs + a * CELL_DIMS + c, sh_q[0][c] + sh_q[1][c] + sh_q[2][c] + sh_q[3][c]);
  }
}

__global__ void cell_centres_kernel(const double *__restrict__ sums, const int *__restrict__ cnt,
                                    float *__restrict__ cen) {
  const int a = blockIdx.x, c = threadIdx.x;  // 64 threads
  cen[a * CELL_DIMS + c] = cnt[a] > 0 ? static_cast<float>(sums[a * CELL_DIMS + c] / cnt[a]) : 0.0f;
}

// pooled within-cell variance per coordinate = (total sum of squares - between-cell part) / n, floored
__global__ void cell_var_kernel(const double *__restrict__ sqs, const float *__restrict__ cen,
                                const int *__restrict__ cnt, int P, int dsub, int64_t n, float *__restrict__ var) {
  __shared__ float sh[CELL_DIMS];
  const int c = threadIdx.x;  // 64 threads
  double tot = 0.0;
  for (int a = 0; a < P; ++a) {
    const double m = cen[a * CELL_DIMS + c];
    tot += sqs[a * CELL_DIMS + c] - static_cast<double>(cnt[a]) * m * m;
  }
  float v = (c < dsub) ? static_cast<float>(fmax(tot, 0.0) / static_cast<double>(n)) : 0.0f;
  sh[c] = v;
  __syncthreads();
  float vmax = 0.0f;
  for (int i = 0; i < CELL_DIMS; ++i) vmax = fmaxf(vmax, sh[i]);
  if (!(vmax > 0.0f)) vmax = 1.0f;
  v = fmaxf(v, 1e-6f * vmax);
  var[c] = (c < dsub) ? v : 1.0f;
}

__global__ void cell_cw_kernel(const float *__restrict__ cen, const float *__restrict__ var, float *__restrict__ cw) {
  const int a = blockIdx.x, c = threadIdx.x;  // 64 threads
  cw[a * CELL_DIMS + c] = cen[a * CELL_DIMS + c] / var[c];
}

// per ordered pair (a, b): the direction w_ab = cw_b - cw_a (bitwise the negative of w_ba; the extent kernel
// forms it the same way, so gap and extents refer to exactly the same vector, whatever it was rounded to),
// 1 / |w_ab|, the centre gap along it, the centre distance; resets the extents
__global__ void __launch_bounds__(128) cell_pairs_kernel(const float *__restrict__ cen, const float *__restrict__ cw,
                                                        int P, float *__restrict__ inv_norm, float *__restrict__ gap,
                                                        float *__restrict__ cdist, unsigned int *__restrict__ ext,
                                                        unsigned int *__restrict__ rad2) {
  __shared__ float ca[CELL_DIMS], wa[CELL_DIMS];
  const int a = blockIdx.x;
  if (threadIdx.x < CELL_DIMS) {
    ca[threadIdx.x] = cen[a * CELL_DIMS + threadIdx.x];
    wa[threadIdx.x] = cw[a * CELL_DIMS + threadIdx.x];
  }
  if (threadIdx.x == 0) rad2[a] = f2ord(0.0f);
  __syncthreads();
  for (int b = threadIdx.x; b < P; b += blockDim.x) {
    float nn = 0.0f, g = 0.0f, dd = 0.0f;
    for (int c = 0; c < CELL_DIMS; ++c) {
      const float dl = cen[b * CELL_DIMS + c] - ca[c];
      const float w = cw[b * CELL_DIMS + c] - wa[c];
      nn = fmaf(w, w, nn);
      g = fmaf(dl, w, g);
      dd = fmaf(dl, dl, dd);
    }
    const float nrm = sqrtf(nn);
    const size_t o = static_cast<size_t>(a) * P + b;
    inv_norm[o] = nrm > 0.0f ? 1.0f / nrm : 0.0f;
    gap[o] = nrm > 0.0f ? g / nrm : 0.0f;
    cdist[o] = sqrtf(dd);
    ext[o] = f2ord(-__int_as_float(0x7f800000));
  }
}

// one CTA (64 threads) per group of 128 permuted rows of cell a, two rows per thread (for the reason given at
// cell_assign_kernel): for every other cell b the largest projection of (x - c_a) on the direction a -> b, and the
// largest squared distance to the centre
__global__ void __launch_bounds__(64) cell_extent_kernel(const float *__restrict__ x, int d, int dsub,
                                                        const int *__restrict__ dims,
                                                        const float *__restrict__ mu, const int *__restrict__ perm,
                                                        const int *__restrict__ off, int P,
                                                        const int *__restrict__ cell_of, const float *__restrict__ cen,
                                                        const float *__restrict__ cw,
                                                        const float *__restrict__ inv_norm,
                                                        unsigned int *__restrict__ ext,
                                                        unsigned int *__restrict__ rad2, int part, int parts) {
  __shared__ __align__(16) float cwb[128][CELL_DIMS];
  __shared__ unsigned int emax[128];
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * CELL_GROUP;
  if (p0 >= off[P]) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const int a = cell_of[perm[p0]];
  if (a % parts != part) return;  // another rank owns this cell; extents and radii are max-reduced
  const int ra = perm[p0 + tid], rb = perm[p0 + 64 + tid];
  unsigned long long xa[CELL_DIMS / 2], xb[CELL_DIMS / 2];
  float r2a = 0.0f, r2b = 0.0f;
#pragma unroll
  for (int c = 0; c < CELL_DIMS; c += 2) {
    float va[2], vb[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool on = c + e < dsub;
      const int col = on ? __ldg(dims + c + e) : 0;
      const float m = on ? __ldg(mu + col) : 0.0f, ce = __ldg(cen + a * CELL_DIMS + c + e);
      va[e] = (on && ra >= 0) ? (__ldg(x + static_cast<int64_t>(ra) * d + col) - m) - ce : 0.0f;
      vb[e] = (on && rb >= 0) ? (__ldg(x + static_cast<int64_t>(rb) * d + col) - m) - ce : 0.0f;
      r2a = fmaf(va[e], va[e], r2a);
      r2b = fmaf(vb[e], vb[e], r2b);
    }
    xa[c / 2] = pack_f32x2(va[0], va[1]);
    xb[c / 2] = pack_f32x2(vb[0], vb[1]);
  }
  {
    unsigned int m = __reduce_max_sync(FULL, f2ord(fmaxf(r2a, r2b)));
    if (lane == 0) atomicMax(rad2 + a, m);
  }
  const unsigned int neg_inf = f2ord(-__int_as_float(0x7f800000));
  for (int b0 = 0; b0 < P; b0 += 128) {
    __syncthreads();
    const int nb = min(128, P - b0);
    for (int i = tid; i < nb * CELL_DIMS; i += 64)
      cwb[i / CELL_DIMS][i % CELL_DIMS] = cw[b0 * CELL_DIMS + i] - __ldg(cw + a * CELL_DIMS + i % CELL_DIMS);  // w_ab
    emax[tid] = neg_inf;
    emax[64 + tid] = neg_inf;
    __syncthreads();
    for (int bl = 0; bl < nb; ++bl) {
      unsigned long long a01 = 0ull, a23 = 0ull, b01 = 0ull, b23 = 0ull;  // four chains per row, two per accumulator
      const ulonglong2 *pv = reinterpret_cast<const ulonglong2 *>(cwb[bl]);
#pragma unroll
      for (int c4 = 0; c4 < CELL_DIMS / 4; ++c4) {
        const ulonglong2 q = pv[c4];
        fma_f32x2(a01, xa[2 * c4 + 0], q.x);
        fma_f32x2(a23, xa[2 * c4 + 1], q.y);
        fma_f32x2(b01, xb[2 * c4 + 0], q.x);
        fma_f32x2(b23, xb[2 * c4 + 1], q.y);
      }
      const float inr = __ldg(inv_norm + static_cast<size_t>(a) * P + b0 + bl);
      float t0, t1, t2, t3;
      unpack_f32x2(a01, t0, t1);
      unpack_f32x2(a23, t2, t3);
      const float proj_a = ((t0 + t1) + (t2 + t3)) * inr;
      unpack_f32x2(b01, t0, t1);
      unpack_f32x2(b23, t2, t3);
      const float proj_b = ((t0 + t1) + (t2 + t3)) * inr;
      const unsigned int oa = (ra >= 0) ? f2ord(proj_a) : neg_inf, ob = (rb >= 0) ? f2ord(proj_b) : neg_inf;
      const unsigned int m = __reduce_max_sync(FULL, max(oa, ob));
      if (lane == 0) atomicMax(&emax[bl], m);
    }
    __syncthreads();
    for (int t = tid; t < nb; t += 64)
      if (emax[t] != neg_inf) atomicMax(ext + static_cast<size_t>(a) * P + b0 + t, emax[t]);
  }
}

// ------------------------------------------------------------------ register-tiled variants of the two O(n P 64) kernels
// cell_assign_kernel / cell_extent_kernel above read every table value from shared memory once per TWO rows: a
// warp-wide 16-byte read occupies the shared-memory pipe for four cycles and feeds four packed FMAs, which bounds them at
// half the FP32 rate (measured 0.96 / 1.35 ms at 1M x 50, P = 256: a third).  Here a CTA of 256 threads multiplies a
// tile of 128 rows x 128 table entries x 64 coordinates SGEMM-style: both operands lie k-major in shared memory, a
// thread owns 8 rows x 8 entries (64 accumulators as 32 packed pairs) and reads 8 + 8 values per coordinate for 64
// FMAs -- a quarter of the shared-memory bytes per FMA.  The sums are single left-to-right chains over the 64
// coordinates (the kernels above use four interleaved chains); neither result needs a particular rounding: the
// partition may be ANY partition, and the extents' FP32 evaluation error is covered ~50x by cell_lb_kernel's margin.
// A rank of a multi-GPU build runs the same code on its rows / cells, so all ranks still agree bit for bit.
constexpr int TL_LD = 128;                                        // rows / table entries per tile
constexpr int TL_SMEM_BYTES = 2 * CELL_DIMS * TL_LD * 4;          // As[64][128] + Bs[64][128]

// acc[i][jp]: row (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4), entries (jp < 2 ? tx * 4 + 2 jp : 64 + tx * 4 + 2 (jp - 2)) + {0, 1}
__device__ __forceinline__ void tile_fma(const float *__restrict__ As, const float *__restrict__ Bs, int ty, int tx,
                                         unsigned long long (&acc)[8][4]) {
#pragma unroll 4
  for (int k = 0; k < CELL_DIMS; ++k) {
    const float4 a0 = *reinterpret_cast<const float4 *>(As + k * TL_LD + ty * 4);
    const float4 a1 = *reinterpret_cast<const float4 *>(As + k * TL_LD + 64 + ty * 4);
    const ulonglong2 b0 = *reinterpret_cast<const ulonglong2 *>(Bs + k * TL_LD + tx * 4);
    const ulonglong2 b1 = *reinterpret_cast<const ulonglong2 *>(Bs + k * TL_LD + 64 + tx * 4);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const unsigned long long ad = pack_f32x2(av[i], av[i]);
      fma_f32x2(acc[i][0], ad, b0.x);
      fma_f32x2(acc[i][1], ad, b0.y);
      fma_f32x2(acc[i][2], ad, b1.x);
      fma_f32x2(acc[i][3], ad, b1.y);
    }
  }
}
__device__ __forceinline__ int tile_row(int ty, int i) { return i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4; }
__device__ __forceinline__ int tile_col(int tx, int jp, int e) {
  return (jp < 2 ? tx * 4 + 2 * jp : 64 + tx * 4 + 2 * (jp - 2)) + e;
}
// table rows [b0, b0 + nb) (64 floats each), minus `sub` (64 floats, or nullptr) -> Bs[k][b]; threads t0, t0 + nt, ...
__device__ __forceinline__ void tile_fill_table(float *Bs, const float *__restrict__ table, int b0, int nb,
                                                const float *__restrict__ sub, int t0, int nt) {
  for (int i4 = t0; i4 < TL_LD * (CELL_DIMS / 4); i4 += nt) {
    const int b = i4 & (TL_LD - 1), k4 = i4 >> 7;
    float4 w = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (b < nb) {
      w = __ldg(reinterpret_cast<const float4 *>(table + static_cast<size_t>(b0 + b) * CELL_DIMS) + k4);
      if (sub) {
        const float4 u = __ldg(reinterpret_cast<const float4 *>(sub) + k4);
        w.x -= u.x;
        w.y -= u.y;
        w.z -= u.z;
        w.w -= u.w;
      }
    }
    Bs[(4 * k4 + 0) * TL_LD + b] = w.x;
    Bs[(4 * k4 + 1) * TL_LD + b] = w.y;
    Bs[(4 * k4 + 2) * TL_LD + b] = w.z;
    Bs[(4 * k4 + 3) * TL_LD + b] = w.w;
  }
}

// nearest pivot of 128 rows per CTA (ties -> lowest pivot)
__global__ void __launch_bounds__(256, 2) cell_assign_tiled_kernel(const float *__restrict__ x, int64_t row0, int64_t n, int d,
                                                                  int dsub, const int *__restrict__ dims,
                                                                  const float *__restrict__ mu, int P,
                                                                  const float *__restrict__ pivots,
                                                                  int *__restrict__ cell_of, int *__restrict__ cnt) {
  extern __shared__ __align__(16) float tl_smem[];
  float *As = tl_smem, *Bs = tl_smem + CELL_DIMS * TL_LD;
  __shared__ float pn[TL_LD];
  __shared__ int hist[CELL_MAX];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int i = tid; i < P; i += 256) hist[i] = 0;
  const int64_t row_base = row0 + static_cast<int64_t>(blockIdx.x) * TL_LD;
  if (tid < TL_LD) {
    const int64_t row = row_base + tid;
#pragma unroll 8
    for (int c = 0; c < CELL_DIMS; ++c) {
      const bool on = c < dsub;
      const int col = on ? __ldg(dims + c) : 0;
      const float m = on ? __ldg(mu + col) : 0.0f;
      As[c * TL_LD + tid] = (on && row < n) ? __ldg(x + row * d + col) - m : 0.0f;
    }
  } else {  // the other half of the CTA stages the first block of pivots meanwhile
    tile_fill_table(Bs, pivots, 0, min(TL_LD, P), nullptr, tid - TL_LD, 128);
  }
  float best[8];
  int best_p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) best[i] = __int_as_float(0x7f800000), best_p[i] = 0;
  for (int p0 = 0; p0 < P; p0 += TL_LD) {
    const int np = min(TL_LD, P - p0);
    __syncthreads();
    if (p0 > 0) tile_fill_table(Bs, pivots, p0, np, nullptr, tid, 256);
    __syncthreads();
    if (tid < TL_LD) {
      float sq = 0.0f;
      for (int c = 0; c < CELL_DIMS; ++c) sq = fmaf(Bs[c * TL_LD + tid], Bs[c * TL_LD + tid], sq);
      pn[tid] = sq;
    }
    __syncthreads();
    unsigned long long acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;
    tile_fma(As, Bs, ty, tx, acc);
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {  // columns in increasing order: a strict < keeps the lowest pivot among equals
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = tile_col(tx, jp, e);
        const float pnc = pn[col];
        if (col < np) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float lo, hi;
            unpack_f32x2(acc[i][jp], lo, hi);
            const float score = fmaf(-2.0f, e ? hi : lo, pnc);
            if (score < best[i]) {
              best[i] = score;
              best_p[i] = p0 + col;
            }
          }
        }
      }
    }
  }
  // the 16 threads that share a row (same ty: the lanes of a half-warp) agree on (score, pivot), lowest pivot among equals
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(FULL, best[i], o);
      const int op = __shfl_xor_sync(FULL, best_p[i], o);
      if (os < best[i] || (os == best[i] && op < best_p[i])) {
        best[i] = os;
        best_p[i] = op;
      }
    }
    const int64_t row = row_base + tile_row(ty, i);
    if (tx == 0 && row < n) {
      cell_of[row] = best_p[i];
      atomicAdd(&hist[best_p[i]], 1);
    }
  }
  __syncthreads();
  for (int i = tid; i < P; i += 256)
    if (hist[i]) atomicAdd(cnt + i, hist[i]);
}

// one CTA per group of 128 permuted rows of cell a: for every other cell b the largest projection of (x - c_a) on the
// direction a -> b, and the largest squared distance to the centre (same quantities as cell_extent_kernel)
__global__ void __launch_bounds__(256, 2) cell_extent_tiled_kernel(const float *__restrict__ x, int d, int dsub,
                                                                  const int *__restrict__ dims,
                                                                  const float *__restrict__ mu, const int *__restrict__ perm,
                                                                  const int *__restrict__ off, int P,
                                                                  const int *__restrict__ cell_of,
                                                                  const float *__restrict__ cen, const float *__restrict__ cw,
                                                                  const float *__restrict__ inv_norm,
                                                                  unsigned int *__restrict__ ext,
                                                                  unsigned int *__restrict__ rad2, int part, int parts) {
  extern __shared__ __align__(16) float tl_smem[];
  float *As = tl_smem, *Bs = tl_smem + CELL_DIMS * TL_LD;
  __shared__ unsigned int emax[TL_LD];
  __shared__ int row_ok[TL_LD];
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * CELL_GROUP;
  if (p0 >= off[P]) return;
  const int tid = threadIdx.x, lane = tid & 31, tx = tid & 15, ty = tid >> 4;
  const int a = cell_of[perm[p0]];
  if (a % parts != part) return;  // another rank owns this cell; extents and radii are max-reduced
  if (tid < TL_LD) {
    const int r = perm[p0 + tid];
    row_ok[tid] = r >= 0;
    float r2 = 0.0f;
#pragma unroll 8
    for (int c = 0; c < CELL_DIMS; ++c) {
      const bool on = c < dsub;
      const int col = on ? __ldg(dims + c) : 0;
      const float m = on ? __ldg(mu + col) : 0.0f, ce = __ldg(cen + a * CELL_DIMS + c);
      const float v = (on && r >= 0) ? (__ldg(x + static_cast<int64_t>(r) * d + col) - m) - ce : 0.0f;
      r2 = fmaf(v, v, r2);
      As[c * TL_LD + tid] = v;
    }
    const unsigned int m = __reduce_max_sync(FULL, f2ord(r2));
    if (lane == 0) atomicMax(rad2 + a, m);
  } else {  // the other half of the CTA stages the first block of directions meanwhile
    // w_ab = cw_b - cw_a, formed exactly as cell_pairs_kernel forms it
    tile_fill_table(Bs, cw, 0, min(TL_LD, P), cw + a * CELL_DIMS, tid - TL_LD, 128);
  }
  const unsigned int neg_inf = f2ord(-__int_as_float(0x7f800000));
  const float ninf = -__int_as_float(0x7f800000);
  for (int b0 = 0; b0 < P; b0 += TL_LD) {
    const int nb = min(TL_LD, P - b0);
    __syncthreads();
    if (b0 > 0) tile_fill_table(Bs, cw, b0, nb, cw + a * CELL_DIMS, tid, 256);
    if (tid < TL_LD) emax[tid] = neg_inf;
    __syncthreads();
    unsigned long long acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;
    tile_fma(As, Bs, ty, tx, acc);
    bool ok[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ok[i] = row_ok[tile_row(ty, i)] != 0;
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float m = ninf;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float lo, hi;
          unpack_f32x2(acc[i][jp], lo, hi);
          if (ok[i]) m = fmaxf(m, e ? hi : lo);
        }
        m = fmaxf(m, __shfl_xor_sync(FULL, m, 16));  // the warp's other eight rows of the same entries
        const int col = tile_col(tx, jp, e);
        // (inv_norm >= 0 and rounding is monotone: the largest projection is the largest sum times inv_norm)
        if (lane < 16 && col < nb && m > ninf)
          atomicMax(&emax[col], f2ord(m * __ldg(inv_norm + static_cast<size_t>(a) * P + b0 + col)));
      }
    }
    __syncthreads();
    if (tid < nb && emax[tid] != neg_inf) atomicMax(ext + static_cast<size_t>(a) * P + b0 + tid, emax[tid]);
  }
}

// ---- pivot assignment on the tensor cores (TF32).  The partition may be ANY partition (the bounds are computed from the
// rows each cell actually received, and the sweep is exact for whatever they are), so the nearest-pivot search needs no
// FP32 arithmetic: mma.sync.m16n8k8 with 10-bit mantissas moves the odd row that sits between two pivots, nothing else.
// CTA = 8 warps x 32 rows; the pivot block (256 x 64, rounded to TF32) lies in shared memory with a row stride of 68
// floats (conflict-free B-fragment reads); a warp keeps its rows' A fragments (2 m-tiles x 8 k-steps) in registers.
// FP32 FFMA on this chip issues at 64 lanes per clock and SM, packed or not (both FP32 variants above measure
// 0.91-0.96 ms at 1M x 50, P = 256 = exactly that rate): this kernel is what lifts the assignment off that ceiling.
constexpr int AM_ROWS = 256;     // rows per CTA
constexpr int AM_PIV = 256;      // pivots per shared-memory block
constexpr int AM_LD = CELL_DIMS + 4;
constexpr int AM_SMEM_BYTES = AM_PIV * AM_LD * 4 + AM_PIV * 4;

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 2) cell_assign_mma_kernel(const float *__restrict__ x, int64_t row0, int64_t n, int d,
                                                                int dsub, const int *__restrict__ dims,
                                                                const float *__restrict__ mu, int P,
                                                                const float *__restrict__ pivots,
                                                                int *__restrict__ cell_of, int *__restrict__ cnt) {
  extern __shared__ __align__(16) float am_smem[];
  float *Ps = am_smem;                   // [AM_PIV][AM_LD], TF32 bit patterns
  float *pn = am_smem + AM_PIV * AM_LD;  // [AM_PIV] squared norms of the rounded pivots
  __shared__ int hist[CELL_MAX];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
  for (int i = tid; i < P; i += 256) hist[i] = 0;
  const int64_t wbase = row0 + static_cast<int64_t>(blockIdx.x) * AM_ROWS + warp * 32;
  // A fragments: m-tile mt, k-step ks: a0 (row gid, col tig) a1 (row gid + 8, col tig) a2 (gid, tig + 4) a3 (gid + 8, tig + 4)
  uint32_t af[2][8][4];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = ks * 8 + tig + 4 * h;
      const bool on = c < dsub;
      const int col = on ? __ldg(dims + c) : 0;
      const float m = on ? __ldg(mu + col) : 0.0f;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int64_t r0 = wbase + mt * 16 + gid, r1 = r0 + 8;
        af[mt][ks][2 * h + 0] = to_tf32((on && r0 < n) ? __ldg(x + r0 * d + col) - m : 0.0f);
        af[mt][ks][2 * h + 1] = to_tf32((on && r1 < n) ? __ldg(x + r1 * d + col) - m : 0.0f);
      }
    }
  }
  float best[2][2];
  int best_p[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) best[mt][h] = __int_as_float(0x7f800000), best_p[mt][h] = 0;
  for (int p0 = 0; p0 < P; p0 += AM_PIV) {
    const int np = min(AM_PIV, P - p0);
    __syncthreads();
    for (int i4 = tid; i4 < AM_PIV * (CELL_DIMS / 4); i4 += 256) {
      const int pl = i4 >> 4, k4 = i4 & 15;
      float4 w = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (pl < np) w = __ldg(reinterpret_cast<const float4 *>(pivots + static_cast<size_t>(p0 + pl) * CELL_DIMS) + k4);
      uint4 t;
      t.x = to_tf32(w.x);
      t.y = to_tf32(w.y);
      t.z = to_tf32(w.z);
      t.w = to_tf32(w.w);
      *reinterpret_cast<uint4 *>(Ps + pl * AM_LD + 4 * k4) = t;
    }
    __syncthreads();
    {
      float sq = 0.0f;
      for (int c = 0; c < CELL_DIMS; ++c) sq = fmaf(Ps[tid * AM_LD + c], Ps[tid * AM_LD + c], sq);
      pn[tid] = sq;
    }
    __syncthreads();
    for (int nt = 0; nt < (np + 7) / 8; ++nt) {
      float acc[2][4] = {{0.0f, 0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f, 0.0f}};
      const float *prow = Ps + (nt * 8 + gid) * AM_LD + tig;  // B fragment: b0 (k tig, n gid), b1 (k tig + 4, n gid)
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t b0 = __float_as_uint(prow[ks * 8]), b1 = __float_as_uint(prow[ks * 8 + 4]);
        mma_tf32(acc[0], af[0][ks], b0, b1);
        mma_tf32(acc[1], af[1][ks], b0, b1);
      }
      // c0 (row gid, col 2 tig) c1 (gid, 2 tig + 1) c2 (gid + 8, 2 tig) c3 (gid + 8, 2 tig + 1); increasing pivot order
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int pl = nt * 8 + 2 * tig + e;
        const float pnc = pn[pl];
        if (pl < np) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float score = fmaf(-2.0f, acc[mt][2 * h + e], pnc);
              if (score < best[mt][h]) {
                best[mt][h] = score;
                best_p[mt][h] = p0 + pl;
              }
            }
        }
      }
    }
  }
  // the four lanes that share a row agree on (score, pivot), lowest pivot among equals
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int o = 1; o < 4; o <<= 1) {
        const float os = __shfl_xor_sync(FULL, best[mt][h], o);
        const int op = __shfl_xor_sync(FULL, best_p[mt][h], o);
        if (os < best[mt][h] || (os == best[mt][h] && op < best_p[mt][h])) {
          best[mt][h] = os;
          best_p[mt][h] = op;
        }
      }
      const int64_t row = wbase + mt * 16 + h * 8 + gid;
      if (tig == 0 && row < n) {
        cell_of[row] = best_p[mt][h];
        atomicAdd(&hist[best_p[mt][h]], 1);
      }
    }
  __syncthreads();
  for (int i = tid; i < P; i += 256)
    if (hist[i]) atomicAdd(cnt + i, hist[i]);
}

static bool cells_use_tiled() {
  const char *e = getenv("VVB_CELLS_TILED");  // 0 = the two-rows-per-thread kernels
  return !(e && atoi(e) == 0);
}
static int cells_assign_mode() {
  const char *e = getenv("VVB_CELLS_ASSIGN");  // 2 = TF32 tensor cores (default), 1 = register-tiled FP32, 0 = two rows per thread
  if (e) return atoi(e);
  return cells_use_tiled() ? 2 : 0;
}
static int cells_tiled_attr() {
  static bool done = false;
  if (!done) {
    VVB_CUDA(cudaFuncSetAttribute(cell_assign_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AM_SMEM_BYTES));
    VVB_CUDA(cudaFuncSetAttribute(cell_assign_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TL_SMEM_BYTES));
    VVB_CUDA(cudaFuncSetAttribute(cell_extent_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TL_SMEM_BYTES));
    done = true;
  }
  return VVB_OK;
}

// lb(a, b) = scale * max(discriminant bound, ball bound) - margin; +inf for empty b, -inf on the diagonal
__global__ void __launch_bounds__(128) cell_lb_kernel(const int *__restrict__ cnt, int P, const float *__restrict__ gap,
                                                     const float *__restrict__ cdist,
                                                     const unsigned int *__restrict__ ext,
                                                     const unsigned int *__restrict__ rad2,
                                                     const float *__restrict__ scale_ymax, float *__restrict__ lb) {
  const int a = blockIdx.x;
  const float s = scale_ymax[0], ymax = scale_ymax[1];
  const float inf = __int_as_float(0x7f800000);
  const float ra = sqrtf(ord2f(rad2[a])) * 1.00001f;
  for (int b = threadIdx.x; b < P; b += blockDim.x) {
    const size_t o = static_cast<size_t>(a) * P + b, ot = static_cast<size_t>(b) * P + a;
    float v;
    if (cnt[b] == 0) {
      v = inf;
    } else if (a == b || cnt[a] == 0) {
      v = -inf;
    } else {
      const float rb = sqrtf(ord2f(rad2[b])) * 1.00001f;
      const float fisher = gap[o] - ord2f(ext[o]) - ord2f(ext[ot]);
      const float ball = cdist[o] - ra - rb;
      // FP32 evaluation error: ~64 x 2^-24 relative to (centre distance + radii) for the dot products and
      // norms, ~2^-23 of the largest centred norm for the centring subtractions; both covered ~50x
      const float margin = 2e-4f * (cdist[o] + ra + rb) * s + 2e-5f * ymax;
      v = fmaxf(fisher, ball) * s - margin;
    }
    lb[o] = v;
  }
}

// The plan in four phases.  A single-GPU build runs them back to back (part 0 of 1).  A multi-GPU build splits the two
// O(n P) kernels -- pivot assignment by ROWS, extents by CELLS (cell a belongs to part a % parts) -- and exchanges three
// small tables between the phases (knn_tensor.cu: knn_tensor_plan_phase says which):
//   phase 0: pivots, cell_of for rows [row_lo, row_hi)                       -> all-gather cell_of
//   phase 1: sizes, offsets, perm (all rows); moments of the owned cells     -> all-reduce SUM sums, sqs
//   phase 2: centres, pooled variance, directions, extents of owned cells    -> all-reduce MAX ext, rad2
//   phase 3: lower bounds
// Every rank ends with bit-identical centres / directions / extents, so the bounds mean the same on all of them.
int cells_build_phase(const CellPlan &L, const float *x, int d, const float *mu, const int *dims, const float *scale_dev,
                      int phase, int part, int parts, int64_t row_lo, int64_t row_hi, cudaStream_t st) {
  const int P = L.P;
  const int64_t n = L.n;
  const unsigned groups = static_cast<unsigned>(L.n_perm_cap / CELL_GROUP);
  if (phase == 0) {
    // the pivot block borrows the centres' buffer (P x CELL_DIMS floats), which is only written in phase 2
    cell_pivots_kernel<<<P, CELL_DIMS, 0, st>>>(x, n, d, L.dsub, dims, mu, P, L.cen);
    VVB_CHECK_LAUNCH();
    VVB_CUDA(cudaMemsetAsync(L.cnt, 0, sizeof(int) * P, st));
    if (row_hi > row_lo && cells_assign_mode() == 2) {
      if (int rc = cells_tiled_attr()) return rc;
      cell_assign_mma_kernel<<<static_cast<unsigned>((row_hi - row_lo + AM_ROWS - 1) / AM_ROWS), 256, AM_SMEM_BYTES, st>>>(
          x, row_lo, row_hi, d, L.dsub, dims, mu, P, L.cen, L.cell_of, L.cnt);
      VVB_CHECK_LAUNCH();
    } else if (row_hi > row_lo && cells_assign_mode() == 1) {
      if (int rc = cells_tiled_attr()) return rc;
      cell_assign_tiled_kernel<<<static_cast<unsigned>((row_hi - row_lo + TL_LD - 1) / TL_LD), 256, TL_SMEM_BYTES, st>>>(
          x, row_lo, row_hi, d, L.dsub, dims, mu, P, L.cen, L.cell_of, L.cnt);
      VVB_CHECK_LAUNCH();
    } else if (row_hi > row_lo) {
      cell_assign_kernel<<<static_cast<unsigned>((row_hi - row_lo + ASSIGN_ROWS - 1) / ASSIGN_ROWS), 128, 0, st>>>(
          x, row_lo, row_hi, d, L.dsub, dims, mu, P, L.cen, L.cell_of, L.cnt);
      VVB_CHECK_LAUNCH();
    }
  } else if (phase == 1) {
    VVB_CUDA(cudaMemsetAsync(L.cnt, 0, sizeof(int) * P, st));
    VVB_CUDA(cudaMemsetAsync(L.perm, 0xFF, sizeof(int) * static_cast<size_t>(L.n_perm_cap), st));
    VVB_CUDA(cudaMemsetAsync(L.sums, 0, sizeof(double) * static_cast<size_t>(P) * CELL_DIMS, st));
    VVB_CUDA(cudaMemsetAsync(L.sqs, 0, sizeof(double) * static_cast<size_t>(P) * CELL_DIMS, st));
    cell_hist_kernel<<<static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 1184)), 256, 0, st>>>(L.cell_of, 0, n,
                                                                                                     nullptr, P, L.cnt);
    VVB_CHECK_LAUNCH();
    cell_scan_kernel<<<1, 1024, 0, st>>>(L.cnt, P, CELL_GROUP, L.off, L.cursor);
    VVB_CHECK_LAUNCH();
    cell_scatter_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(L.cell_of, 0, n, nullptr, L.cursor,
                                                                                L.perm, -1);
    VVB_CHECK_LAUNCH();
    cell_moments_kernel<<<groups, 256, 0, st>>>(x, d, L.dsub, dims, mu, L.perm, L.off, P, L.cell_of, L.sums, L.sqs, part,
                                                parts);
    VVB_CHECK_LAUNCH();
  } else if (phase == 2) {
    cell_centres_kernel<<<P, CELL_DIMS, 0, st>>>(L.sums, L.cnt, L.cen);
    VVB_CHECK_LAUNCH();
    cell_var_kernel<<<1, CELL_DIMS, 0, st>>>(L.sqs, L.cen, L.cnt, P, L.dsub, n, L.var);
    VVB_CHECK_LAUNCH();
    cell_cw_kernel<<<P, CELL_DIMS, 0, st>>>(L.cen, L.var, L.cw);
    VVB_CHECK_LAUNCH();
    cell_pairs_kernel<<<P, 128, 0, st>>>(L.cen, L.cw, P, L.inv_norm, L.gap, L.cdist, L.ext, L.rad2);
    VVB_CHECK_LAUNCH();
    if (cells_use_tiled()) {
      if (int rc = cells_tiled_attr()) return rc;
      cell_extent_tiled_kernel<<<groups, 256, TL_SMEM_BYTES, st>>>(x, d, L.dsub, dims, mu, L.perm, L.off, P, L.cell_of, L.cen,
                                                                    L.cw, L.inv_norm, L.ext, L.rad2, part, parts);
    } else {
      cell_extent_kernel<<<groups, 64, 0, st>>>(x, d, L.dsub, dims, mu, L.perm, L.off, P, L.cell_of, L.cen, L.cw,
                                                 L.inv_norm, L.ext, L.rad2, part, parts);
    }
    VVB_CHECK_LAUNCH();
  } else {
    cell_lb_kernel<<<P, 128, 0, st>>>(L.cnt, P, L.gap, L.cdist, L.ext, L.rad2, scale_dev, L.lb);
    VVB_CHECK_LAUNCH();
  }
  return VVB_OK;
}

int cells_build(const CellPlan &L, const float *x, int d, const float *mu, const int *dims, const float *scale_dev,
                cudaStream_t st) {
  for (int phase = 0; phase < 4; ++phase) {
    const int rc = cells_build_phase(L, x, d, mu, dims, scale_dev, phase, 0, 1, 0, L.n, st);
    if (rc) return rc;
  }
  return VVB_OK;
}

// Cell-sorted list of the query rows [q0, q0+nq) -- or, with `row_list`, of the nq original rows it names.
int cells_query_list(const CellPlan &L, int64_t q0, int64_t nq, int64_t nq_pad, int *qlist, const int64_t *row_list,
                     cudaStream_t st) {
  VVB_CUDA(cudaMemsetAsync(L.qcnt, 0, sizeof(int) * L.P, st));
  cell_hist_kernel<<<static_cast<unsigned>(std::min<int64_t>((nq + 255) / 256, 1184)), 256, 0, st>>>(
      L.cell_of, q0, nq, row_list, L.P, L.qcnt);
  VVB_CHECK_LAUNCH();
  cell_scan_kernel<<<1, 1024, 0, st>>>(L.qcnt, L.P, 1, L.qoff, L.qcursor);
  VVB_CHECK_LAUNCH();
  cell_scatter_kernel<<<static_cast<unsigned>((nq_pad + 255) / 256), 256, 0, st>>>(L.cell_of, q0, nq, row_list, L.qcursor,
                                                                                   qlist, nq_pad);
  VVB_CHECK_LAUNCH();
  return VVB_OK;
}


// ------------------------------------------------------------------ longest-first order of the query blocks
// A query block's sweep is roughly as long as its own neighbourhood: the tiles of the cells its bound cannot
// separate from it (bound <= 0).  Blocks are launched in decreasing order of that estimate, so the long ones start
// first and the short ones fill the tail of the launch (a few hundred blocks of uneven length over 148 SMs otherwise
// end with SMs idling behind the longest block: the multi-GPU shards are that small).
constexpr int CTA_BUCKETS = 1024;

// one warp per block of `rows_per_block` query-list entries: estimate -> bucket histogram
__global__ void __launch_bounds__(256) cta_estimate_kernel(const int *__restrict__ qlist, int64_t nq, int rows_per_block,
                                                          int n_blocks, const int *__restrict__ cell_of,
                                                          const int *__restrict__ off, const float *__restrict__ lb, int P,
                                                          int *__restrict__ est, int *__restrict__ hist) {
  const int lane = threadIdx.x & 31;
  const int blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (blk >= n_blocks) return;
  const int64_t first = static_cast<int64_t>(blk) * rows_per_block;
  const int64_t last = min(first + rows_per_block, nq) - 1;
  const int a0 = cell_of[qlist[first]], a1 = cell_of[qlist[last]];  // the list is sorted by cell
  int tiles = 0;
  for (int b = lane; b < P; b += 32) {
    float v = __int_as_float(0x7f800000);
    for (int a = a0; a <= a1; ++a) v = fminf(v, __ldg(lb + static_cast<size_t>(a) * P + b));
    if (v <= 0.0f) tiles += (off[b + 1] - off[b]) / CELL_GROUP;
  }
  tiles = __reduce_add_sync(FULL, tiles);
  if (lane == 0) {
    const int total = off[P] / CELL_GROUP;
    const int bucket = min(CTA_BUCKETS - 1, static_cast<int>(static_cast<int64_t>(tiles) * CTA_BUCKETS / max(1, total)));
    est[blk] = bucket;
    atomicAdd(hist + bucket, 1);
  }
}
// descending exclusive scan of the bucket histogram (largest estimates first)
__global__ void __launch_bounds__(CTA_BUCKETS) cta_scan_kernel(const int *__restrict__ hist, int *__restrict__ cursor) {
  __shared__ int sh[CTA_BUCKETS];
  const int tid = threadIdx.x;
  const int mine = hist[CTA_BUCKETS - 1 - tid];
  sh[tid] = mine;
  __syncthreads();
  for (int o = 1; o < CTA_BUCKETS; o <<= 1) {
    const int v = (tid >= o) ? sh[tid - o] : 0;
    __syncthreads();
    sh[tid] += v;
    __syncthreads();
  }
  cursor[CTA_BUCKETS - 1 - tid] = sh[tid] - mine;
}
__global__ void __launch_bounds__(256) cta_scatter_kernel(const int *__restrict__ est, int n_blocks, int *__restrict__ cursor,
                                                         int *__restrict__ order) {
  const int blk = blockIdx.x * blockDim.x + threadIdx.x;
  if (blk < n_blocks) order[atomicAdd(cursor + est[blk], 1)] = blk;
}

size_t cells_cta_order_scratch_ints(int64_t n_blocks) { return static_cast<size_t>(n_blocks) + 2 * CTA_BUCKETS; }

// order[i] = the query block launched i-th; scratch: cells_cta_order_scratch_ints(n_blocks) ints
int cells_cta_order(const CellPlan &L, const int *qlist, int64_t nq, int rows_per_block, int n_blocks, int *scratch,
                    int *order, cudaStream_t st) {
  int *est = scratch, *hist = scratch + n_blocks, *cursor = hist + CTA_BUCKETS;
  VVB_CUDA(cudaMemsetAsync(hist, 0, sizeof(int) * CTA_BUCKETS, st));
  cta_estimate_kernel<<<(n_blocks + 7) / 8, 256, 0, st>>>(qlist, nq, rows_per_block, n_blocks, L.cell_of, L.off, L.lb, L.P,
                                                         est, hist);
  VVB_CHECK_LAUNCH();
  cta_scan_kernel<<<1, CTA_BUCKETS, 0, st>>>(hist, cursor);
  VVB_CHECK_LAUNCH();
  cta_scatter_kernel<<<(n_blocks + 255) / 256, 256, 0, st>>>(est, n_blocks, cursor, order);
  VVB_CHECK_LAUNCH();
  return VVB_OK;
}

}  // namespace vvb
