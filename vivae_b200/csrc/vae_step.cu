// One training step of the ViVAE autoencoder as ONE kernel: forward, the three loss terms, backward.
//
// Replaces, inside `ViVAE.train_epoch` (/root/reference/src/vivae/model.py:130-177), the forward of
// `Autoencoder.forward` (network.py:219-296: encode -> reparameterise -> KL -> decode -> MSE -> quartet loss) and
// its autograd backward -- about 140 kernels of a few microseconds each on a 32k-parameter network, i.e. a step
// bound by launch count, not by arithmetic -- for the default configuration (variational, exact-erf GELU, euclidean
// quartet distances, no geometric / imitation terms).  The network stays an ordinary nn.Module: the kernel reads the
// Linear layers' parameter tensors in place and leaves the gradient of every parameter in one flat buffer that the
// parameters' `.grad` fields view, so `torch.optim.Adam` (model.py:256) steps as before.
//
// Work split: a CTA owns 8 quartets = 32 rows of the batch (rows {q, nq+q, 2nq+q, 3nq+q}, losses.py:326-336), so
// the quartet term is local to it.  Activations live in shared memory TRANSPOSED ([feature][row], row stride 36
// floats): lane = row, so a layer is `acc += W[o][i] * act[i][lane]` with the weight a warp-uniform (broadcast) read
// of the layer's weight matrix -- staged into shared memory right before the layer runs (at most 32 KB at a time:
// reading the parameter tensors through L1 left every multiply waiting on an L2 round trip, 290 us per step against
// 40 us now) -- and the activation a conflict-free shared-memory read.
// Only pre-activations are kept; GELU and its derivative are recomputed where needed (a 128 x 32 tile at most).
// Backward runs in place: dU_l overwrites U_l.  Weight gradients: one (o, i) entry per lane and 32-row dot product,
// written to a per-CTA partial buffer (no atomics: the step is run-to-run deterministic, VIVAE_DETERMINISTIC);
// `vae_reduce_kernel` adds the partials in CTA order and finishes the three loss terms.
#include <algorithm>

#include "common.cuh"

namespace vvb {

constexpr int VS_R = 32;          // rows per CTA
constexpr int VS_RS = 36;         // row stride of the transposed activation buffers (16-byte aligned, conflict-free)
#ifndef VVB_VS_THREADS
#define VVB_VS_THREADS 512
#endif
constexpr int VS_THREADS = VVB_VS_THREADS;  // 16 warps: the layers are chains of dependent FMAs behind shared-memory reads
constexpr int VS_WARPS = VS_THREADS / 32;
constexpr int VS_MAX_LAYERS = 16;

struct VsLayer {
  const float *W, *b;  // (out, in) row-major, (out)
  int in, out;
  int goff;            // offset of dW in the flat gradient; db follows at goff + in * out
  int ubuf;            // shared-memory offset (floats) of this layer's output (pre-activation) buffer
};
struct VsParams {
  VsLayer layer[VS_MAX_LAYERS];  // encoder layers, mu head, logvar head, decoder layers
  int n_enc, n_dec;
  int B, D, L, nq, P;
  int off_x, off_eps, off_z, off_dzq, off_t, off_w, off_w2;  // shared-memory offsets (floats); two weight buffers
  float lam_recon, lam_kldiv, lam_mds;
  const float *x, *eps;
  float *gpart;  // (gridDim.x, P)
  float *lpart;  // (gridDim.x, 4): sum of the KL integrand, sum of squared errors, sum of quartet costs
};

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}

// T[i][r] = gelu(U[i][r])
__device__ __forceinline__ void act_into(float *T, const float *U, int dim) {
  for (int idx = threadIdx.x; idx < dim * VS_R; idx += VS_THREADS) {
    const int i = idx >> 5, r = idx & 31;
    T[i * VS_RS + r] = gelu_f(U[i * VS_RS + r]);
  }
}

// A layer's weight matrix (out x in, row-major) and bias into shared memory, with cp.async (round 2c): the weights of
// the NEXT layer travel into the other of two buffers while the current layer is computed -- the step is a chain of
// ~22 short phases on 16 SMs, and every phase used to begin with an exposed L2 round trip for its weights (a plain
// load + store staging loop in front of a barrier).
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void stage_weights_async(const VsLayer &ly, float *Wsm) {
  const int nw = ly.in * ly.out;
  const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(Wsm));
  if ((nw % 4 == 0) && ((reinterpret_cast<uintptr_t>(ly.W) & 15) == 0)) {
    for (int i = threadIdx.x; i < nw / 4; i += VS_THREADS) cp_async_16(dst + 16u * i, ly.W + 4 * i);
  } else {
    for (int i = threadIdx.x; i < nw; i += VS_THREADS) cp_async_4(dst + 4u * i, ly.W + i);
  }
  for (int i = threadIdx.x; i < ly.out; i += VS_THREADS) cp_async_4(dst + 4u * (nw + i), ly.b + i);
}
// the order in which the step needs its layers' weights: forward through all layers, then backward through the decoder,
// the two heads, and the encoder down to its second layer (the first layer's backward needs no weights)
__device__ __forceinline__ int vs_seq(const VsParams &p, int si) {
  const int nl = p.n_enc + 2 + p.n_dec;
  if (si < nl) return si;
  si -= nl;
  if (si < p.n_dec) return nl - 1 - si;
  si -= p.n_dec;
  if (si < 2) return p.n_enc + si;
  si -= 2;
  if (si < p.n_enc - 1) return p.n_enc - 1 - si;
  return -1;
}

// out[o][lane] = b[o] + sum_i W[o][i] * in[i][lane]; a warp computes four outputs at a time.  Wsm: staged weights.
__device__ __forceinline__ void gemm_fwd(const VsLayer &ly, const float *Wsm, const float *in, float *out, int warp,
                                         int lane) {
  const int I = ly.in, O = ly.out;
  const bool vec = (I % 4 == 0);
  const float *bsm = Wsm + I * O;
  for (int ob = warp * 4; ob < O; ob += VS_WARPS * 4) {
    float acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = (ob + j < O) ? bsm[ob + j] : 0.0f;
    if (vec) {
      for (int i = 0; i < I; i += 4) {
        const float a0 = in[(i + 0) * VS_RS + lane], a1 = in[(i + 1) * VS_RS + lane];
        const float a2 = in[(i + 2) * VS_RS + lane], a3 = in[(i + 3) * VS_RS + lane];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (ob + j < O) {
            const float4 w = *reinterpret_cast<const float4 *>(Wsm + (ob + j) * I + i);
            acc[j] = fmaf(w.x, a0, acc[j]);
            acc[j] = fmaf(w.y, a1, acc[j]);
            acc[j] = fmaf(w.z, a2, acc[j]);
            acc[j] = fmaf(w.w, a3, acc[j]);
          }
        }
      }
    } else {
      for (int i = 0; i < I; ++i) {
        const float a = in[i * VS_RS + lane];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ob + j < O) acc[j] = fmaf(Wsm[(ob + j) * I + i], a, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (ob + j < O) out[(ob + j) * VS_RS + lane] = acc[j];
  }
}

// dW[o][i] = sum_r dU[o][r] * A[i][r],  db[o] = sum_r dU[o][r]   -> this CTA's partial gradient
__device__ __forceinline__ void gemm_dw(const VsLayer &ly, const float *dU, const float *A, float *g, int warp, int lane) {
  const int I = ly.in, O = ly.out;
  for (int o = warp; o < O; o += VS_WARPS) {
    float4 du[8];
    const float4 *dr = reinterpret_cast<const float4 *>(dU + o * VS_RS);
#pragma unroll
    for (int v = 0; v < 8; ++v) du[v] = dr[v];  // broadcast reads: every lane holds the whole row
    for (int ib = 0; ib < I; ib += 32) {
      const int i = ib + lane;
      if (i < I) {
        const float4 *ar = reinterpret_cast<const float4 *>(A + i * VS_RS);
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;  // four chains: the FMA latency, not the rate, bounds one
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const float4 a = ar[v];
          s0 = fmaf(du[v].x, a.x, s0);
          s1 = fmaf(du[v].y, a.y, s1);
          s2 = fmaf(du[v].z, a.z, s2);
          s3 = fmaf(du[v].w, a.w, s3);
        }
        g[ly.goff + o * I + i] = (s0 + s1) + (s2 + s3);
      }
    }
    if (lane == 0) {
      float s = 0.0f;
#pragma unroll
      for (int v = 0; v < 8; ++v) s += (du[v].x + du[v].y) + (du[v].z + du[v].w);
      g[ly.goff + I * O + o] = s;
    }
  }
}

// acc[i][lane] (+)= sum_o W[o][i] * dU[o][lane]; four inputs i at a time per warp.
//   MODE 0: Uprev[i][lane] = acc * gelu'(Uprev[i][lane])      (back through the activation, in place)
//   MODE 1: dst[i][lane] += acc                                (first decoder layer: into dZ)
//   MODE 2: dst[i][lane]  = acc                                (raw, no activation: used for the first head)
template <int MODE>
__device__ __forceinline__ void gemm_dx(const VsLayer &ly, const float *Wsm, const float *dU, float *dst, int warp,
                                        int lane) {
  const int I = ly.in, O = ly.out;
  const bool vec = (I % 4 == 0);
  for (int ib = warp * 4; ib < I; ib += VS_WARPS * 4) {
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
    for (int o = 0; o < O; ++o) {
      const float gq = dU[o * VS_RS + lane];
      if (vec) {
        const float4 w = *reinterpret_cast<const float4 *>(Wsm + o * I + ib);
        acc[0] = fmaf(w.x, gq, acc[0]);
        acc[1] = fmaf(w.y, gq, acc[1]);
        acc[2] = fmaf(w.z, gq, acc[2]);
        acc[3] = fmaf(w.w, gq, acc[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ib + j < I) acc[j] = fmaf(Wsm[o * I + ib + j], gq, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = ib + j;
      if (i < I) {
        float *p = dst + i * VS_RS + lane;
        if (MODE == 0) *p = acc[j] * gelu_grad_f(*p);
        else if (MODE == 1) *p += acc[j];
        else *p = acc[j];
      }
    }
  }
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

__global__ void __launch_bounds__(VS_THREADS, 1) vae_step_kernel(const __grid_constant__ VsParams p) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[3][VS_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = p.B, D = p.D, L = p.L, nq = p.nq;
  const int n_layers = p.n_enc + 2 + p.n_dec;
  float *X = sm + p.off_x, *EPS = sm + p.off_eps, *Z = sm + p.off_z, *DZQ = sm + p.off_dzq, *T = sm + p.off_t;
  const VsLayer &ly_mu = p.layer[p.n_enc], &ly_lv = p.layer[p.n_enc + 1];
  float *MU = sm + ly_mu.ubuf, *LV = sm + ly_lv.ubuf;
  float *XH = sm + p.layer[n_layers - 1].ubuf;
  float *g = p.gpart + static_cast<size_t>(blockIdx.x) * p.P;
  // local row r = a * 8 + ql  <->  batch row a * nq + 8 * blockIdx.x + ql   (quartet 8 * blockIdx.x + ql, member a)
  const int q0 = blockIdx.x * 8;

  // ---- weights: double-buffered, prefetched one layer ahead.  acquire(li) replaces the synchronous staging of layer li
  // (the caller's __syncthreads follows it); prefetch() right after that barrier starts the next layer's copy into the
  // buffer the PREVIOUS phase read, which every thread has left by then.
  float *const Wbuf0 = sm + p.off_w, *const Wbuf1 = sm + p.off_w2;
  int w_si = 0, w_cur = 0;
  stage_weights_async(p.layer[0], Wbuf0);
  auto acquire = [&](int li) -> float * {
    if (vs_seq(p, w_si) != li) __trap();  // the call order and vs_seq must agree
    cp_async_wait_all();
    return w_cur ? Wbuf1 : Wbuf0;
  };
  auto prefetch = [&]() {
    ++w_si;
    w_cur ^= 1;
    const int nx = vs_seq(p, w_si);
    if (nx >= 0) stage_weights_async(p.layer[nx], w_cur ? Wbuf1 : Wbuf0);
  };

  // ---- load the batch rows (transposed); rows of quartets beyond nq are zero and take no part
  for (int idx = threadIdx.x; idx < VS_R * D; idx += VS_THREADS) {
    const int r = idx / D, c = idx - r * D;
    const int q = q0 + (r & 7);
    X[c * VS_RS + r] = (q < nq) ? __ldg(p.x + (static_cast<size_t>(r >> 3) * nq + q) * D + c) : 0.0f;
  }
  for (int idx = threadIdx.x; idx < VS_R * L; idx += VS_THREADS) {
    const int r = idx / L, c = idx - r * L;
    const int q = q0 + (r & 7);
    EPS[c * VS_RS + r] = (q < nq) ? __ldg(p.eps + (static_cast<size_t>(r >> 3) * nq + q) * L + c) : 0.0f;
  }
  __syncthreads();

  // ---- encoder (every layer: activation of the previous one into T, weights into Wsm, then the product)
  float *Wsm;
  for (int l = 0; l < p.n_enc; ++l) {
    const float *in = X;
    if (l > 0) {
      act_into(T, sm + p.layer[l - 1].ubuf, p.layer[l - 1].out);
      in = T;
    }
    Wsm = acquire(l);
    __syncthreads();
    prefetch();
    gemm_fwd(p.layer[l], Wsm, in, sm + p.layer[l].ubuf, warp, lane);
    __syncthreads();
  }
  act_into(T, sm + p.layer[p.n_enc - 1].ubuf, p.layer[p.n_enc - 1].out);
  Wsm = acquire(p.n_enc);
  __syncthreads();
  prefetch();
  gemm_fwd(ly_mu, Wsm, T, MU, warp, lane);
  __syncthreads();
  Wsm = acquire(p.n_enc + 1);
  __syncthreads();
  prefetch();
  gemm_fwd(ly_lv, Wsm, T, LV, warp, lane);
  __syncthreads();
  // z = mu + eps * exp(logvar / 2)   (network.py:113-130)
  for (int idx = threadIdx.x; idx < L * VS_R; idx += VS_THREADS) {
    const int c = idx >> 5, r = idx & 31;
    Z[c * VS_RS + r] = MU[c * VS_RS + r] + EPS[c * VS_RS + r] * expf(0.5f * LV[c * VS_RS + r]);
  }
  __syncthreads();
  // ---- decoder
  for (int l = 0; l < p.n_dec; ++l) {
    const int li = p.n_enc + 2 + l;
    const float *in = Z;
    if (l > 0) {
      act_into(T, sm + p.layer[li - 1].ubuf, p.layer[li - 1].out);
      in = T;
    }
    Wsm = acquire(li);
    __syncthreads();
    prefetch();
    gemm_fwd(p.layer[li], Wsm, in, sm + p.layer[li].ubuf, warp, lane);
    __syncthreads();
  }

  // ---- loss terms and the gradients they start the backward pass with
  float kl_sum = 0.0f, se_sum = 0.0f, cost_sum = 0.0f;
  {
    const float eps_sq = 1e-4f, log_eps_sq = -9.210340371976182f;  // KLDivLoss: eps_std = 1e-2 (losses.py:105-131)
    for (int idx = threadIdx.x; idx < L * VS_R; idx += VS_THREADS) {
      const int c = idx >> 5, r = idx & 31;
      if (q0 + (r & 7) < nq) {
        const float mu = MU[c * VS_RS + r], lv = LV[c * VS_RS + r];
        kl_sum += lv + log_eps_sq - mu * mu - eps_sq * expf(lv);
      }
    }
    const float gs = p.lam_recon * 2.0f / (static_cast<float>(B) * static_cast<float>(D));
    for (int idx = threadIdx.x; idx < D * VS_R; idx += VS_THREADS) {
      const int c = idx >> 5, r = idx & 31;
      float diff = 0.0f;
      if (q0 + (r & 7) < nq) diff = XH[c * VS_RS + r] - X[c * VS_RS + r];
      se_sum = fmaf(diff, diff, se_sum);
      XH[c * VS_RS + r] = gs * diff;  // d(lam_recon * MSE) / d xhat
    }
  }
  {
    // quartet term (losses.py:237-264): warp ql owns quartet q0 + ql
    constexpr int pa[6] = {0, 1, 2, 0, 0, 1};
    constexpr int pb[6] = {1, 2, 3, 2, 3, 3};
    const int ql = warp & 7;
    const bool mine = warp < 8;  // eight quartets, sixteen warps
    float sx[6] = {0, 0, 0, 0, 0, 0}, sz[6] = {0, 0, 0, 0, 0, 0};
    for (int c = lane; c < D; c += 32) {
      float v[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) v[a] = X[c * VS_RS + a * 8 + ql];
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        const float t = v[pa[e]] - v[pb[e]];
        sx[e] = fmaf(t, t, sx[e]);
      }
    }
    for (int c = lane; c < L; c += 32) {
      float v[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) v[a] = Z[c * VS_RS + a * 8 + ql];
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        const float t = v[pa[e]] - v[pb[e]];
        sz[e] = fmaf(t, t, sz[e]);
      }
    }
    float dx[6], dz[6], rdz[6], Xs = 0.0f, Ss = 0.0f;
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      sx[e] = warp_sum_f(sx[e]);
      sz[e] = warp_sum_f(sz[e]);
      dx[e] = sqrtf(fmaxf(sx[e], 1e-9f));
      const float cz = fmaxf(sz[e], 1e-9f);
      rdz[e] = rsqrtf(cz);
      dz[e] = cz * rdz[e];
      Xs += dx[e];
      Ss += dz[e];
    }
    const float invX = 1.0f / Xs, invS = 1.0f / Ss;
    float ee[6], cost = 0.0f, er = 0.0f;
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      const float nz = dz[e] * invS;
      ee[e] = nz - dx[e] * invX;
      cost = fmaf(ee[e], ee[e], cost);
      er = fmaf(ee[e], nz, er);
    }
    const bool valid = mine && q0 + ql < nq;
    if (valid && lane == 0) cost_sum += cost;
    // d(lam_mds * mean_q(cost) * B / nq) / dz
    const float gsc = p.lam_mds * static_cast<float>(B) / (static_cast<float>(nq) * static_cast<float>(nq));
    for (int c = lane; c < L; c += 32) {
      float v[4], gz[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int a = 0; a < 4; ++a) v[a] = Z[c * VS_RS + a * 8 + ql];
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        if (sz[e] > 1e-9f) {
          const float t = 2.0f * invS * (ee[e] - er) * gsc * (v[pa[e]] - v[pb[e]]) * rdz[e];
          gz[pa[e]] += t;
          gz[pb[e]] -= t;
        }
      }
      if (mine) {
#pragma unroll
        for (int a = 0; a < 4; ++a) DZQ[c * VS_RS + a * 8 + ql] = valid ? gz[a] : 0.0f;
      }
    }
  }
  kl_sum = warp_sum_f(kl_sum);
  se_sum = warp_sum_f(se_sum);
  cost_sum = warp_sum_f(cost_sum);
  if (lane == 0) {
    red[0][warp] = kl_sum;
    red[1][warp] = se_sum;
    red[2][warp] = cost_sum;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.0f;
    for (int w = 0; w < VS_WARPS; ++w) s += red[threadIdx.x][w];
    p.lpart[blockIdx.x * 4 + threadIdx.x] = s;
  }

  // ---- backward: decoder (dU of the last layer is in XH), in place
  for (int l = p.n_dec - 1; l >= 0; --l) {
    const int li = p.n_enc + 2 + l;
    const float *A = Z;
    if (l > 0) {
      act_into(T, sm + p.layer[li - 1].ubuf, p.layer[li - 1].out);
      A = T;
    }
    Wsm = acquire(li);
    __syncthreads();
    prefetch();
    const float *dU = sm + p.layer[li].ubuf;
    gemm_dw(p.layer[li], dU, A, g, warp, lane);
    if (l > 0) gemm_dx<0>(p.layer[li], Wsm, dU, sm + p.layer[li - 1].ubuf, warp, lane);
    else gemm_dx<1>(p.layer[li], Wsm, dU, DZQ, warp, lane);  // dZ = quartet gradient + decoder gradient
    __syncthreads();
  }
  // ---- through the reparameterisation and the KL term: dMU, dLV overwrite MU, LV
  {
    const float kq = p.lam_kldiv / (static_cast<float>(B) * static_cast<float>(B) * static_cast<float>(L));
    const float eps_sq = 1e-4f;
    for (int idx = threadIdx.x; idx < L * VS_R; idx += VS_THREADS) {
      const int c = idx >> 5, r = idx & 31;
      const bool valid = q0 + (r & 7) < nq;
      const float mu = MU[c * VS_RS + r], lv = LV[c * VS_RS + r], dzv = DZQ[c * VS_RS + r];
      const float sd = expf(0.5f * lv);
      MU[c * VS_RS + r] = valid ? dzv + kq * mu : 0.0f;
      LV[c * VS_RS + r] = valid ? dzv * EPS[c * VS_RS + r] * 0.5f * sd - 0.5f * kq * (1.0f - eps_sq * expf(lv)) : 0.0f;
    }
  }
  act_into(T, sm + p.layer[p.n_enc - 1].ubuf, p.layer[p.n_enc - 1].out);
  Wsm = acquire(p.n_enc);
  __syncthreads();
  prefetch();
  gemm_dw(ly_mu, MU, T, g, warp, lane);
  gemm_dw(ly_lv, LV, T, g, warp, lane);
  {
    // gradient of the last hidden activation through both heads; X (dead since the loss terms, and as tall as the
    // heads' input) is the accumulator, each (i, row) entry owned by the same thread in both calls
    float *ACC = X;
    gemm_dx<2>(ly_mu, Wsm, MU, ACC, warp, lane);
    __syncthreads();
    Wsm = acquire(p.n_enc + 1);
    __syncthreads();
    prefetch();
    gemm_dx<1>(ly_lv, Wsm, LV, ACC, warp, lane);
    __syncthreads();
    float *U = sm + p.layer[p.n_enc - 1].ubuf;
    for (int idx = threadIdx.x; idx < p.layer[p.n_enc - 1].out * VS_R; idx += VS_THREADS) {
      const int i = idx >> 5, r = idx & 31;
      U[i * VS_RS + r] = ACC[i * VS_RS + r] * gelu_grad_f(U[i * VS_RS + r]);
    }
  }
  __syncthreads();
  // ---- backward: encoder.  X was used as scratch: reload it for the first layer's weight gradient
  for (int l = p.n_enc - 1; l >= 0; --l) {
    const float *A;
    if (l > 0) {
      act_into(T, sm + p.layer[l - 1].ubuf, p.layer[l - 1].out);
      Wsm = acquire(l);
      A = T;
    } else {
      for (int idx = threadIdx.x; idx < VS_R * D; idx += VS_THREADS) {
        const int r = idx / D, c = idx - r * D;
        const int q = q0 + (r & 7);
        X[c * VS_RS + r] = (q < nq) ? __ldg(p.x + (static_cast<size_t>(r >> 3) * nq + q) * D + c) : 0.0f;
      }
      A = X;
    }
    __syncthreads();
    if (l > 0) prefetch();
    const float *dU = sm + p.layer[l].ubuf;
    gemm_dw(p.layer[l], dU, A, g, warp, lane);
    if (l > 0) gemm_dx<0>(p.layer[l], Wsm, dU, sm + p.layer[l - 1].ubuf, warp, lane);
    __syncthreads();
  }
}

// Sum of the per-CTA partial gradients in CTA order (deterministic) + the three weighted loss terms, added to the
// epoch's running sums (slots of model._TERMS: recon 0, kldiv 1, mds 4) and written to loss_out[3].
__global__ void __launch_bounds__(256) vae_reduce_kernel(const float *__restrict__ gpart, int ncta, int P,
                                                        float *__restrict__ gflat, const float *__restrict__ lpart,
                                                        int B, int D, int L, int nq, float lam_recon, float lam_kldiv,
                                                        float lam_mds, float *__restrict__ sums,
                                                        float *__restrict__ loss_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) {
    float s = 0.0f;
    for (int c = 0; c < ncta; ++c) s += gpart[static_cast<size_t>(c) * P + i];
    gflat[i] = s;
  }
  if (i == 0) {
    float kl = 0.0f, se = 0.0f, co = 0.0f;
    for (int c = 0; c < ncta; ++c) {
      kl += lpart[c * 4 + 0];
      se += lpart[c * 4 + 1];
      co += lpart[c * 4 + 2];
    }
    const float recon = lam_recon * se / (static_cast<float>(B) * static_cast<float>(D));
    const float kldiv = lam_kldiv * (-0.5f * kl / (static_cast<float>(B) * static_cast<float>(L))) / static_cast<float>(B);
    const float mds = lam_mds * (co / static_cast<float>(nq)) * static_cast<float>(B) / static_cast<float>(nq);
    loss_out[0] = recon;
    loss_out[1] = kldiv;
    loss_out[2] = mds;
    if (sums) {
      sums[0] += recon;
      sums[1] += kldiv;
      sums[4] += mds;
    }
  }
}

// Host side.  dims: layer l maps in_dims[l] -> out_dims[l]; layers = n_enc encoder layers, mu, logvar, n_dec decoder
// layers.  Returns the dynamic shared memory the kernel needs through *smem_bytes (call with gpart == nullptr to ask).
int launch_vae_step(const float *const *w_ptrs, const float *const *b_ptrs, const int *in_dims, const int *out_dims,
                    int n_enc, int n_dec, const float *x, const float *eps, int B, int D, int L, float lam_recon,
                    float lam_kldiv, float lam_mds, float *gpart, float *lpart, float *gflat, float *sums,
                    float *loss_out, size_t *smem_bytes, int *n_params, int *n_cta, cudaStream_t st) {
  const int n_layers = n_enc + 2 + n_dec;
  VVB_SUPPORTED(n_enc >= 1 && n_dec >= 1 && n_layers <= VS_MAX_LAYERS, "fused step: unsupported layer count");
  VVB_SUPPORTED(B >= 4 && B % 4 == 0 && L >= 1 && D >= 1, "fused step: batch size must be a positive multiple of 4");
  VsParams p;
  int off = 0, P = 0, tmax = 0;
  auto take = [&](int dim) {
    const int o = off;
    off += dim * VS_RS;
    return o;
  };
  p.off_x = take(D > out_dims[n_enc - 1] ? D : out_dims[n_enc - 1]);  // also the scratch of the heads' backward
  p.off_eps = take(L);
  p.off_z = take(L);
  p.off_dzq = take(L);
  for (int l = 0; l < n_layers; ++l) {
    VsLayer &ly = p.layer[l];
    ly.W = w_ptrs ? w_ptrs[l] : nullptr;
    ly.b = b_ptrs ? b_ptrs[l] : nullptr;
    ly.in = in_dims[l];
    ly.out = out_dims[l];
    ly.goff = P;
    P += ly.in * ly.out + ly.out;
    ly.ubuf = take(ly.out);
    if (ly.out > tmax) tmax = ly.out;
  }
  p.off_t = take(tmax);
  {
    int wmax = 0;  // staging buffer of one layer's weights + bias
    for (int l = 0; l < n_layers; ++l) wmax = std::max(wmax, in_dims[l] * out_dims[l] + out_dims[l]);
    off = (off + 3) / 4 * 4;
    p.off_w = off;
    off += (wmax + 3) / 4 * 4;
    p.off_w2 = off;  // second buffer: the next layer's weights arrive (cp.async) while this layer is computed
    off += (wmax + 3) / 4 * 4;
  }
  // shape checks: the chain must be consistent
  bool ok = in_dims[0] == D && out_dims[n_enc] == L && out_dims[n_enc + 1] == L && in_dims[n_enc] == out_dims[n_enc - 1] &&
            in_dims[n_enc + 1] == out_dims[n_enc - 1] && in_dims[n_enc + 2] == L && out_dims[n_layers - 1] == D;
  for (int l = 1; l < n_enc; ++l) ok = ok && in_dims[l] == out_dims[l - 1];
  for (int l = 1; l < n_dec; ++l) ok = ok && in_dims[n_enc + 2 + l] == out_dims[n_enc + 2 + l - 1];
  VVB_REQUIRE(ok, "fused step: layer dimensions do not chain");
  const size_t smem = static_cast<size_t>(off) * sizeof(float);
  const int nq = B / 4;
  const int ncta = (nq + 7) / 8;
  if (smem_bytes) *smem_bytes = smem;
  if (n_params) *n_params = P;
  if (n_cta) *n_cta = ncta;
  VVB_SUPPORTED(smem <= 227u * 1024u, "fused step: the activations of 32 rows need %zu bytes of shared memory", smem);
  if (!gpart) return VVB_OK;  // size query
  p.n_enc = n_enc;
  p.n_dec = n_dec;
  p.B = B;
  p.D = D;
  p.L = L;
  p.nq = nq;
  p.P = P;
  p.lam_recon = lam_recon;
  p.lam_kldiv = lam_kldiv;
  p.lam_mds = lam_mds;
  p.x = x;
  p.eps = eps;
  p.gpart = gpart;
  p.lpart = lpart;
  static size_t attr_set = 0;
  if (smem > attr_set) {
    VVB_CUDA(cudaFuncSetAttribute(vae_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set = smem;
  }
  vae_step_kernel<<<ncta, VS_THREADS, smem, st>>>(p);
  VVB_CHECK_LAUNCH();
  vae_reduce_kernel<<<(P + 255) / 256, 256, 0, st>>>(gpart, ncta, P, gflat, lpart, B, D, L, nq, lam_recon, lam_kldiv,
                                                     lam_mds, sums, loss_out);
  VVB_CHECK_LAUNCH();
  return VVB_OK;
}

}  // namespace vvb
