// Quartet ("stochastic MDS") loss, fused forward + gradient.
// Replaces MDSLoss.__call__ (/root/reference/src/vivae/losses.py:266-349), which costs
// ~1.5k ATen dispatches per batch for forward+backward.
//
// EIGHT lanes per quartet q of a sampling pass (four quartets per warp): rows perm[a*nq + q],
// a = 0..3 (losses.py:326,332-336).  The six pairs, in the order of losses.py:234,263, are
// (0,1) (1,2) (2,3) (0,2) (0,3) (1,3).  The lanes of a group stride over the D (and L) coordinates
// with 8-byte loads (D even), partial sums are combined with three xor-shuffles, then every lane
// holds the six high-dimensional and six latent distances and evaluates
//     cost_q = sum_p (dz_p/S - dx_p/X)^2,   X = sum_p dx_p,  S = sum_p dz_p
// and the closed-form gradient
//     dC/d(dz_p) = (2/S) (e_p - sum_p' e_p' dz_p'/S),   e_p = dz_p/S - dx_p/X.
// Each row belongs to exactly one quartet per pass, so gradient rows are written
// without atomics and the result is run-to-run deterministic (VIVAE_DETERMINISTIC).
// CTAs walk the quartets with a grid stride (a grid of SMs x 8 CTAs, 16 row loads in flight per
// lane: bandwidth-sized batches stream at HBM speed); the scalar loss is reduced in a fixed order:
// lane tree -> warp slots -> one partial per CTA -> the last CTA to finish adds the partials.
#include <cstdlib>

#include "common.cuh"

namespace vvb {

constexpr int QT_THREADS = 256;  // 32 quartets per CTA and iteration
constexpr int QT_GROUP = 8;      // lanes per quartet
constexpr int QT_PER_CTA = QT_THREADS / QT_GROUP;

// sum over the 8 lanes of a quartet's group (xor offsets below 8 never leave the aligned group)
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = QT_GROUP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// Reduced pair statistics of four rows: for euclidean the six squared distances, for
// cosine additionally the four squared norms and six dot products.
struct QuadStats {
  float sq[6];   // sum (a-b)^2
  float dot[6];  // sum a*b        (cosine only)
  float nrm[4];  // sum a*a        (cosine only)
};

template <int METRIC>
__device__ __forceinline__ void quad_accum(QuadStats &s, float v0, float v1, float v2, float v3) {
  if (METRIC == VVB_DIST_EUCLIDEAN) {
    float t;
    t = v0 - v1; s.sq[0] = fmaf(t, t, s.sq[0]);
    t = v1 - v2; s.sq[1] = fmaf(t, t, s.sq[1]);
    t = v2 - v3; s.sq[2] = fmaf(t, t, s.sq[2]);
    t = v0 - v2; s.sq[3] = fmaf(t, t, s.sq[3]);
    t = v0 - v3; s.sq[4] = fmaf(t, t, s.sq[4]);
    t = v1 - v3; s.sq[5] = fmaf(t, t, s.sq[5]);
  } else {
    s.nrm[0] = fmaf(v0, v0, s.nrm[0]); s.nrm[1] = fmaf(v1, v1, s.nrm[1]);
    s.nrm[2] = fmaf(v2, v2, s.nrm[2]); s.nrm[3] = fmaf(v3, v3, s.nrm[3]);
    s.dot[0] = fmaf(v0, v1, s.dot[0]); s.dot[1] = fmaf(v1, v2, s.dot[1]);
    s.dot[2] = fmaf(v2, v3, s.dot[2]); s.dot[3] = fmaf(v0, v2, s.dot[3]);
    s.dot[4] = fmaf(v0, v3, s.dot[4]); s.dot[5] = fmaf(v1, v3, s.dot[5]);
  }
}

// `sub` = lane within the quartet's group of 8.  VEC2: rows are 8-byte aligned (dim even): float2 loads.
template <int METRIC, bool VEC2>
__device__ __forceinline__ void quad_stats(const float *base, const int64_t (&r)[4], int dim, int sub,
                                           QuadStats &s) {
#pragma unroll
  for (int p = 0; p < 6; ++p) s.sq[p] = 0.f, s.dot[p] = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) s.nrm[a] = 0.f;
  if (VEC2) {
    const float2 *b0 = reinterpret_cast<const float2 *>(base + r[0] * dim);
    const float2 *b1 = reinterpret_cast<const float2 *>(base + r[1] * dim);
    const float2 *b2 = reinterpret_cast<const float2 *>(base + r[2] * dim);
    const float2 *b3 = reinterpret_cast<const float2 *>(base + r[3] * dim);
    const int half = dim >> 1;
#pragma unroll 4
    for (int c = sub; c < half; c += QT_GROUP) {
      const float2 v0 = __ldg(b0 + c), v1 = __ldg(b1 + c), v2 = __ldg(b2 + c), v3 = __ldg(b3 + c);
      quad_accum<METRIC>(s, v0.x, v1.x, v2.x, v3.x);
      quad_accum<METRIC>(s, v0.y, v1.y, v2.y, v3.y);
    }
  } else {
    for (int c = sub; c < dim; c += QT_GROUP)
      quad_accum<METRIC>(s, __ldg(base + r[0] * dim + c), __ldg(base + r[1] * dim + c), __ldg(base + r[2] * dim + c),
                         __ldg(base + r[3] * dim + c));
  }
  if (METRIC == VVB_DIST_EUCLIDEAN) {
#pragma unroll
    for (int p = 0; p < 6; ++p) s.sq[p] = group_sum(s.sq[p]);
  } else {
#pragma unroll
    for (int p = 0; p < 6; ++p) s.dot[p] = group_sum(s.dot[p]);
#pragma unroll
    for (int a = 0; a < 4; ++a) s.nrm[a] = group_sum(s.nrm[a]);
  }
}

// dist[p] and, for euclidean, rdist[p] = 1 / dist[p] (one MUFU.RSQ per pair instead of a square root and,
// in the gradient, a division: the kernel is bound by instruction issue as much as by bandwidth)
template <int METRIC>
__device__ __forceinline__ void pair_dists(const QuadStats &s, float (&dist)[6], float (&rdist)[6]) {
  constexpr int pa[6] = {0, 1, 2, 0, 0, 1};
  constexpr int pb[6] = {1, 2, 3, 2, 3, 3};
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    if (METRIC == VVB_DIST_EUCLIDEAN) {
      const float c = fmaxf(s.sq[p], 1e-9f);  // losses.py:154-159: sqrt(max(sum, 1e-9))
      rdist[p] = rsqrtf(c);
      dist[p] = c * rdist[p];
    } else {
      const float ca = fmaxf(sqrtf(s.nrm[pa[p]]), 1e-8f), cb = fmaxf(sqrtf(s.nrm[pb[p]]), 1e-8f);
      dist[p] = 1.0f - s.dot[p] / (ca * cb);  // losses.py:176 (ATen clamps each norm at eps=1e-8)
      rdist[p] = 0.0f;
    }
  }
}

template <int MHD, int MLD, bool VEC2>
__global__ void __launch_bounds__(QT_THREADS) quartet_fwd_kernel(
    const float *__restrict__ x, const float *__restrict__ z, int64_t B, int D, int L,
    const int64_t *__restrict__ perm, float *__restrict__ partials, float *__restrict__ grad_unit,
    int accumulate, float grad_scale, float pass_weight, float *__restrict__ loss_out,
    int accumulate_loss, unsigned int *__restrict__ done_counter) {
  constexpr int pa[6] = {0, 1, 2, 0, 0, 1};
  constexpr int pb[6] = {1, 2, 3, 2, 3, 3};
  const int64_t nq = B / 4;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane & (QT_GROUP - 1);
  float cost_sum = 0.f;  // this lane's quartets (counted by sub == 0 only)

  // rows beyond the last whole quartet take no part: their gradient is zero (first pass only)
  if (grad_unit != nullptr && !accumulate && blockIdx.x == 0)
    for (int64_t i = 4 * nq * L + threadIdx.x; i < B * L; i += QT_THREADS) {
      const int64_t row = i / L;
      grad_unit[(perm ? perm[row] : row) * L + i % L] = 0.f;
    }

  // whole warps iterate together (the shuffles need every lane): q0 = first quartet of the warp's four
  for (int64_t q0 = static_cast<int64_t>(blockIdx.x) * QT_PER_CTA + warp * 4; q0 < nq;
       q0 += static_cast<int64_t>(gridDim.x) * QT_PER_CTA) {
    const int64_t q = q0 + (lane >> 3);
    const bool valid = q < nq;
    int64_t r[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) r[a] = valid ? (perm ? perm[a * nq + q] : a * nq + q) : 0;

    QuadStats sx, sz;
    quad_stats<MHD, VEC2>(x, r, D, sub, sx);
    quad_stats<MLD, false>(z, r, L, sub, sz);
    float dx[6], dz[6], rdx[6], rdz[6];
    pair_dists<MHD>(sx, dx, rdx);
    pair_dists<MLD>(sz, dz, rdz);
    float X = 0.f, S = 0.f;
#pragma unroll
    for (int p = 0; p < 6; ++p) X += dx[p], S += dz[p];
    const float invX = 1.0f / X, invS = 1.0f / S;
    float e[6], cost = 0.f, er = 0.f;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      const float nz = dz[p] * invS;
      e[p] = nz - dx[p] * invX;
      cost = fmaf(e[p], e[p], cost);
      er = fmaf(e[p], nz, er);
    }
    if (valid && sub == 0) cost_sum += cost;

    if (grad_unit != nullptr && valid) {
      float w[6];
#pragma unroll
      for (int p = 0; p < 6; ++p) w[p] = 2.0f * invS * (e[p] - er) * grad_scale;
      for (int c = sub; c < L; c += QT_GROUP) {
        float zv[4], g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int a = 0; a < 4; ++a) zv[a] = __ldg(z + r[a] * L + c);
#pragma unroll
        for (int p = 0; p < 6; ++p) {
          const int a = pa[p], b = pb[p];
          if (MLD == VVB_DIST_EUCLIDEAN) {
            // d sqrt(max(s,1e-9))/dz_a = (z_a - z_b)/dist when s > 1e-9, else 0
            if (sz.sq[p] > 1e-9f) {
              const float t = w[p] * (zv[a] - zv[b]) * rdz[p];
              g[a] += t;
              g[b] -= t;
            }
          } else {
            const float na = sqrtf(sz.nrm[a]), nb = sqrtf(sz.nrm[b]);
            const float ca = fmaxf(na, 1e-8f), cb = fmaxf(nb, 1e-8f);
            const float inv = 1.0f / (ca * cb);
            // cos = dot/(ca cb);  d(1-cos)/dz_a = -(z_b/(ca cb) - dot/(ca^2 cb) * z_a/|a|)
            const float da = na > 0.f ? zv[a] / na : 0.f, db = nb > 0.f ? zv[b] / nb : 0.f;
            g[a] -= w[p] * (zv[b] * inv - sz.dot[p] * inv / ca * da);
            g[b] -= w[p] * (zv[a] * inv - sz.dot[p] * inv / cb * db);
          }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          float *gp = grad_unit + r[a] * L + c;
          *gp = accumulate ? (*gp + g[a]) : g[a];
        }
      }
    }
  }

  // ---- deterministic loss reduction: lane tree -> warp slots -> CTA partial -> last CTA adds the partials
  __shared__ bool is_last;
  __shared__ float red[QT_THREADS];
  cost_sum = warp_sum(cost_sum);
  if (lane == 0) red[warp] = cost_sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < QT_THREADS / 32; ++w) t += red[w];
    partials[blockIdx.x] = t;
    __threadfence();
    const unsigned int ticket = atomicAdd(done_counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float part = 0.f;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += QT_THREADS) part += __ldcg(partials + i);
  red[threadIdx.x] = part;
  __syncthreads();
#pragma unroll
  for (int o = QT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // pass loss = mean_q(cost) * B / nq  (losses.py:263,339-343); pass_weight = 1/n_sampling
    const float pass = red[0] / static_cast<float>(nq) * static_cast<float>(B) / static_cast<float>(nq);
    const float contrib = pass * pass_weight;
    *loss_out = accumulate_loss ? (*loss_out + contrib) : contrib;
    *done_counter = 0;  // ready for the next launch on this stream
  }
}

// Bandwidth-sized batches (round 2c).  The kernel above repeats the per-quartet arithmetic (distances, cost, gradient
// weights: ~550 of its ~750 warp instructions per four quartets) in all eight lanes of a quartet's group, i.e. once per
// FOUR quartets and warp; at B = 2^20 that made it issue-bound (ncu: 66 % issue-active at 44 % occupancy, 0.46 of the HBM
// peak).  Here a warp takes 32 quartets per round: eight load phases of four quartets each (eight lanes stride over a
// quartet's D coordinates exactly as above), after each of which lane j picks up the reduced statistics of quartet j with
// one shuffle per value -- and then ONE pass of the per-quartet arithmetic with lane = quartet, latent rows read and
// gradient rows written by that lane.  Same arithmetic per quartet, 2.8x fewer instructions.  Used when there are enough
// quartets to fill the GPU that way (launch_quartet_fwd); training-size batches keep the kernel above (latency counts
// there, not issue slots).
// (compiled for 4 CTAs per SM = 64 registers, the compiler's own choice; measured at B = 2^20: forcing 5 CTAs (48
// registers, 160 bytes of spills) 84.9 us, 6 CTAs (40 registers) 99.4 us, no bound on the registers 92.2 us, against 70.4)
template <int MHD, int MLD, bool VEC2>
__global__ void __launch_bounds__(QT_THREADS) quartet_fwd_big_kernel(
    const float *__restrict__ x, const float *__restrict__ z, int64_t B, int D, int L,
    const int64_t *__restrict__ perm, float *__restrict__ partials, float *__restrict__ grad_unit,
    int accumulate, float grad_scale, float pass_weight, float *__restrict__ loss_out,
    int accumulate_loss, unsigned int *__restrict__ done_counter) {
  constexpr int pa[6] = {0, 1, 2, 0, 0, 1};
  constexpr int pb[6] = {1, 2, 3, 2, 3, 3};
  const int64_t nq = B / 4;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane & (QT_GROUP - 1), grp = lane >> 3;
  float cost_sum = 0.f;

  if (grad_unit != nullptr && !accumulate && blockIdx.x == 0)
    for (int64_t i = 4 * nq * L + threadIdx.x; i < B * L; i += QT_THREADS) {
      const int64_t row = i / L;
      grad_unit[(perm ? perm[row] : row) * L + i % L] = 0.f;
    }

  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (QT_THREADS / 32);
  for (int64_t Q0 = (static_cast<int64_t>(blockIdx.x) * (QT_THREADS / 32) + warp) * 32; Q0 < nq; Q0 += warps_total * 32) {
    // ---- eight load phases: statistics of quartet Q0 + 4 g + grp, picked up by lane 4 g + grp
    QuadStats sx;
#pragma unroll
    for (int p = 0; p < 6; ++p) sx.sq[p] = 0.f, sx.dot[p] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) sx.nrm[a] = 0.f;
#pragma unroll 2
    for (int g = 0; g < 8; ++g) {
      const int64_t q = Q0 + 4 * g + grp;
      const bool valid = q < nq;
      int64_t r[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) r[a] = valid ? (perm ? perm[a * nq + q] : a * nq + q) : 0;
      QuadStats sg;
      quad_stats<MHD, VEC2>(x, r, D, sub, sg);
      const int src = 8 * (lane & 3);       // any lane of the group that holds quartet Q0 + lane in this phase
      const bool take = (lane >> 2) == g;
      if (MHD == VVB_DIST_EUCLIDEAN) {
#pragma unroll
        for (int p = 0; p < 6; ++p) {
          const float t = __shfl_sync(FULL, sg.sq[p], src);
          if (take) sx.sq[p] = t;
        }
      } else {
#pragma unroll
        for (int p = 0; p < 6; ++p) {
          const float t = __shfl_sync(FULL, sg.dot[p], src);
          if (take) sx.dot[p] = t;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float t = __shfl_sync(FULL, sg.nrm[a], src);
          if (take) sx.nrm[a] = t;
        }
      }
    }
    // ---- lane = quartet
    const int64_t q = Q0 + lane;
    const bool valid = q < nq;
    int64_t r[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) r[a] = valid ? (perm ? perm[a * nq + q] : a * nq + q) : 0;
    QuadStats sz;
#pragma unroll
    for (int p = 0; p < 6; ++p) sz.sq[p] = 0.f, sz.dot[p] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) sz.nrm[a] = 0.f;
    for (int c = 0; c < L; ++c)
      quad_accum<MLD>(sz, __ldg(z + r[0] * L + c), __ldg(z + r[1] * L + c), __ldg(z + r[2] * L + c), __ldg(z + r[3] * L + c));
    float dx[6], dz[6], rdx[6], rdz[6];
    pair_dists<MHD>(sx, dx, rdx);
    pair_dists<MLD>(sz, dz, rdz);
    float X = 0.f, S = 0.f;
#pragma unroll
    for (int p = 0; p < 6; ++p) X += dx[p], S += dz[p];
    const float invX = 1.0f / X, invS = 1.0f / S;
    float e[6], cost = 0.f, er = 0.f;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      const float nz = dz[p] * invS;
      e[p] = nz - dx[p] * invX;
      cost = fmaf(e[p], e[p], cost);
      er = fmaf(e[p], nz, er);
    }
    if (valid) cost_sum += cost;

    if (grad_unit != nullptr && valid) {
      float w[6];
#pragma unroll
      for (int p = 0; p < 6; ++p) w[p] = 2.0f * invS * (e[p] - er) * grad_scale;
      for (int c = 0; c < L; ++c) {
        float zv[4], g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int a = 0; a < 4; ++a) zv[a] = __ldg(z + r[a] * L + c);
#pragma unroll
        for (int p = 0; p < 6; ++p) {
          const int a = pa[p], b = pb[p];
          if (MLD == VVB_DIST_EUCLIDEAN) {
            if (sz.sq[p] > 1e-9f) {
              const float t = w[p] * (zv[a] - zv[b]) * rdz[p];
              g[a] += t;
              g[b] -= t;
            }
          } else {
            const float na = sqrtf(sz.nrm[a]), nb = sqrtf(sz.nrm[b]);
            const float ca = fmaxf(na, 1e-8f), cb = fmaxf(nb, 1e-8f);
            const float inv = 1.0f / (ca * cb);
            const float da = na > 0.f ? zv[a] / na : 0.f, db = nb > 0.f ? zv[b] / nb : 0.f;
            g[a] -= w[p] * (zv[b] * inv - sz.dot[p] * inv / ca * da);
            g[b] -= w[p] * (zv[a] * inv - sz.dot[p] * inv / cb * db);
          }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          float *gp = grad_unit + r[a] * L + c;
          *gp = accumulate ? (*gp + g[a]) : g[a];
        }
      }
    }
  }

  // ---- deterministic loss reduction (as above)
  __shared__ bool is_last;
  __shared__ float red[QT_THREADS];
  cost_sum = warp_sum(cost_sum);
  if (lane == 0) red[warp] = cost_sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < QT_THREADS / 32; ++w) t += red[w];
    partials[blockIdx.x] = t;
    __threadfence();
    const unsigned int ticket = atomicAdd(done_counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float part = 0.f;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += QT_THREADS) part += __ldcg(partials + i);
  red[threadIdx.x] = part;
  __syncthreads();
#pragma unroll
  for (int o = QT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float pass = red[0] / static_cast<float>(nq) * static_cast<float>(B) / static_cast<float>(nq);
    const float contrib = pass * pass_weight;
    *loss_out = accumulate_loss ? (*loss_out + contrib) : contrib;
    *done_counter = 0;
  }
}

__global__ void quartet_nan_kernel(float *loss_out) { *loss_out = __int_as_float(0x7fc00000); }

__global__ void scale_by_scalar_kernel(const float *__restrict__ g, const float *__restrict__ s,
                                       int64_t count, float *__restrict__ out) {
  const float f = __ldg(s);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = g[i] * f;
}

size_t quartet_workspace_bytes(int64_t B) {
  // [done counter (256 B)][cost_q: B/4 floats]
  return 256 + align_up(static_cast<size_t>(B / 4 + 1) * sizeof(float), 256);
}

int launch_quartet_fwd(const float *x, const float *z, int64_t B, int D, int L, const int64_t *const *perms,
                       int n_sampling, int mhd, int mld, float *loss_out, float *grad_unit, void *workspace,
                       cudaStream_t st) {
  VVB_REQUIRE(B >= 0 && D >= 1 && L >= 1 && n_sampling >= 1, "quartet loss: bad shape arguments");
  VVB_REQUIRE((mhd == 0 || mhd == 1) && (mld == 0 || mld == 1), "Invalid distance function specification for MDS loss");
  const int64_t nq = B / 4;
  if (nq == 0) {  // mean over zero quartets: NaN, like the reference; no row gets a gradient
    if (grad_unit && B > 0) VVB_CUDA(cudaMemsetAsync(grad_unit, 0, static_cast<size_t>(B) * L * sizeof(float), st));
    quartet_nan_kernel<<<1, 1, 0, st>>>(loss_out);
    VVB_CHECK_LAUNCH();
    return VVB_OK;
  }
  // workspace: [done counter (256 B), zero between launches: the last CTA resets it][one partial per CTA]
  unsigned int *counter = static_cast<unsigned int *>(workspace);
  float *partials = reinterpret_cast<float *>(static_cast<char *>(workspace) + 256);
  VVB_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
  // grid = the CTAs that are resident at once (a grid-stride loop does the rest: no partial last wave).  Batches with
  // at least one round of 32 quartets for every warp of a full grid take the 32-quartets-per-warp kernel.
  const bool vec2 = (D % 2 == 0) && (reinterpret_cast<uintptr_t>(x) % 8 == 0);
  static int occ_cache[2][2][2][2] = {};  // [big][mhd][mld][vec2]; 0 = not asked yet
  const char *qb = getenv("VVB_QUARTET_BIG");  // 0 / 1 forces the kernel
  auto occupancy = [&](int big) -> int {
    int &occ = occ_cache[big][mhd][mld][vec2 ? 1 : 0];
    if (occ == 0) {
      cudaError_t oe = cudaSuccess;
#define VVB_QT_OCC(K, A, Bm, V) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, K<A, Bm, V>, QT_THREADS, 0)
#define VVB_QT_OCC_ALL(K)                                                                              \
  do {                                                                                                 \
    if (mhd == 0 && mld == 0) { if (vec2) VVB_QT_OCC(K, 0, 0, true); else VVB_QT_OCC(K, 0, 0, false); } \
    else if (mhd == 0) { if (vec2) VVB_QT_OCC(K, 0, 1, true); else VVB_QT_OCC(K, 0, 1, false); }        \
    else if (mld == 0) { if (vec2) VVB_QT_OCC(K, 1, 0, true); else VVB_QT_OCC(K, 1, 0, false); }        \
    else { if (vec2) VVB_QT_OCC(K, 1, 1, true); else VVB_QT_OCC(K, 1, 1, false); }                      \
  } while (0)
      if (big) VVB_QT_OCC_ALL(quartet_fwd_big_kernel);
      else VVB_QT_OCC_ALL(quartet_fwd_kernel);
#undef VVB_QT_OCC_ALL
#undef VVB_QT_OCC
      if (oe != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = 2;
      }
    }
    return occ;
  };
  const int64_t big_min = static_cast<int64_t>(sm_count()) * occupancy(1) * (QT_THREADS / 32) * 32;
  const bool big = qb ? atoi(qb) != 0 : nq >= big_min;
  const int64_t per_cta = big ? (QT_THREADS / 32) * 32 : QT_PER_CTA;
  int64_t grid = (nq + per_cta - 1) / per_cta;
  const int64_t resident = static_cast<int64_t>(sm_count()) * occupancy(big ? 1 : 0);
  if (grid > resident) grid = resident;
  const float grad_scale = static_cast<float>(static_cast<double>(B) / (static_cast<double>(nq) * nq) / n_sampling);
  const float pass_weight = 1.0f / static_cast<float>(n_sampling);
  for (int s = 0; s < n_sampling; ++s) {
    const int64_t *perm = perms ? perms[s] : nullptr;
#define VVB_QT_LAUNCH_K(K, A, Bm)                                                                                   \
  do {                                                                                                              \
    if (vec2)                                                                                                       \
      K<A, Bm, true><<<static_cast<unsigned>(grid), QT_THREADS, 0, st>>>(                                           \
          x, z, B, D, L, perm, partials, grad_unit, s > 0, grad_scale, pass_weight, loss_out, s > 0, counter);      \
    else                                                                                                            \
      K<A, Bm, false><<<static_cast<unsigned>(grid), QT_THREADS, 0, st>>>(                                          \
          x, z, B, D, L, perm, partials, grad_unit, s > 0, grad_scale, pass_weight, loss_out, s > 0, counter);      \
  } while (0)
#define VVB_QT_LAUNCH(A, Bm)                                          \
  do {                                                                \
    if (big) VVB_QT_LAUNCH_K(quartet_fwd_big_kernel, A, Bm);          \
    else VVB_QT_LAUNCH_K(quartet_fwd_kernel, A, Bm);                  \
  } while (0)
    if (mhd == 0 && mld == 0) VVB_QT_LAUNCH(0, 0);
    else if (mhd == 0 && mld == 1) VVB_QT_LAUNCH(0, 1);
    else if (mhd == 1 && mld == 0) VVB_QT_LAUNCH(1, 0);
    else VVB_QT_LAUNCH(1, 1);
#undef VVB_QT_LAUNCH
#undef VVB_QT_LAUNCH_K
    VVB_CHECK_LAUNCH();
  }
  return VVB_OK;
}

int launch_quartet_bwd(const float *grad_unit, const float *grad_out, int64_t count, float *grad_z,
                       cudaStream_t st) {
  if (count <= 0) return VVB_OK;
  int64_t blocks = (count + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  scale_by_scalar_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(grad_unit, grad_out, count, grad_z);
  VVB_CHECK_LAUNCH();
  return VVB_OK;
}

}  // namespace vvb
