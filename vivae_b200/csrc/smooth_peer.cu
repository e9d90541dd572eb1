// Gauss-Seidel kNN smoother over several GPUs in ONE sweep: compute fused with the exchange over NVLink peer memory.
//
// The reference's sweep (/root/reference/src/vivae/knn.py:173-185) is sequential in the rows -- row i reads rows
// j < i already updated -- so it does not split into independent row shards, and running it whole on every GPU
// (round 1) leaves a term that does not shrink with the GPU count.  Here the rows are INTERLEAVED over the ranks
// (row i belongs to rank i mod N) and every rank keeps a complete copy of the sweep's output in memory its peers can
// write (CUDA IPC, NVLink / NVSwitch): the owner of row i computes it from its local copies and stores it into ALL
// ranks' copies (N - 1 remote stores of the row over NVLink: the all-gather, row by row, inside the kernel).  Every
// gather is local; the only remote traffic is the push of finished rows (n (N-1)/N rows out per GPU).
//
// Publication without fences.  A first version published a row with a flag after a system-scope fence and acquired
// it with another: bit-exact, but fence.sc.sys is the scarce resource -- measured on 2 B200 at 1M x 50, one per row
// on either side costs 2.1 ms of a 3.7 ms sweep (1.5 ms with gpu-scope fences, which do not cover a peer), and
// batching them delays the peers more than it saves.  So the flag travels WITH the data (the "LL" idea of collective
// libraries): every element of the output copy is one 64-bit word {value bits, sweep epoch}, written with a single
// 8-byte store -- single-copy atomic, also over NVLink -- and a reader that finds the current epoch in the upper half
// has the value in the lower half, with no ordering between different words required.  A row that needs a neighbour
// still in flight simply re-reads the word (back-off) until the epoch matches.  float64 elements are two such words.
// Deadlock-free for any grid size: every rank hands out its rows in increasing order, so the lowest unfinished row of
// the whole sweep is held by a resident warp and depends on finished rows only.
// The arithmetic is the single-GPU kernel's (sequential sum in neighbour order, true divide, three separately rounded
// operations): the result is bit-identical to it and to the reference.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace vvb {

constexpr int SP_THREADS = 256;
constexpr int SP_UNROLL = 8;
constexpr int SP_MAX_RANKS = 8;

template <typename T> struct ArithP;
template <> struct ArithP<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct ArithP<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

struct PeerPtrs {
  unsigned long long *ll[SP_MAX_RANKS];  // every rank's copy of the sweep's output as {value, epoch} words
};

// One 8-byte access, served by L2 (never L1): L2 is where the peers' stores into this GPU's memory land, so the word is
// either the old one or the new one, whole.  (ld.relaxed.sys gives the same guarantee by the book, but every such load
// is issued as a strong system-scope access that does not pipeline: the sweep ran 3x slower with it -- 8.4 ms against
// 2.6 ms at 1M x 50 on 2 GPUs.)
__device__ __forceinline__ unsigned long long ld_word(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_word(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// words per element: 1 (float) or 2 (double: low half, high half)
template <typename T> struct LL;
template <> struct LL<float> {
  static constexpr int W = 1;
  static __device__ __forceinline__ bool ready(const unsigned long long *w, unsigned epoch) {
    return static_cast<unsigned>(w[0] >> 32) == epoch;
  }
  static __device__ __forceinline__ float value(const unsigned long long *w) {
    return __uint_as_float(static_cast<unsigned>(w[0]));
  }
  static __device__ __forceinline__ void pack(float v, unsigned epoch, unsigned long long *w) {
    w[0] = (static_cast<unsigned long long>(epoch) << 32) | __float_as_uint(v);
  }
};
template <> struct LL<double> {
  static constexpr int W = 2;
  static __device__ __forceinline__ bool ready(const unsigned long long *w, unsigned epoch) {
    return static_cast<unsigned>(w[0] >> 32) == epoch && static_cast<unsigned>(w[1] >> 32) == epoch;
  }
  static __device__ __forceinline__ double value(const unsigned long long *w) {
    return __longlong_as_double(static_cast<long long>((w[1] << 32) | (w[0] & 0xffffffffull)));
  }
  static __device__ __forceinline__ void pack(double v, unsigned epoch, unsigned long long *w) {
    const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
    w[0] = (static_cast<unsigned long long>(epoch) << 32) | (b & 0xffffffffull);
    w[1] = (static_cast<unsigned long long>(epoch) << 32) | (b >> 32);
  }
};

// DPL = columns per lane held in registers, KPL = neighbour indices per lane (k <= 32 * KPL)
template <typename T, int DPL, int KPL>
__global__ void __launch_bounds__(SP_THREADS) smooth_peer_kernel(const T *__restrict__ prev, const PeerPtrs pp,
                                                                const int64_t *__restrict__ idx, int64_t ld_idx,
                                                                int64_t n, int d, int k, T coef, unsigned epoch,
                                                                int rank, int world, unsigned long long *ticket,
                                                                int *bad_index) {
  constexpr int W = LL<T>::W;
  const int lane = threadIdx.x & 31;
  const T kk = static_cast<T>(k);
  const unsigned long long *next = pp.ll[rank];
  for (;;) {
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(ticket, 1ull);
    t = __shfl_sync(FULL, t, 0);
    const int64_t i = static_cast<int64_t>(t) * world + rank;  // this rank's rows, in increasing order
    if (i >= n) break;
    const int i32 = static_cast<int>(i);
    const int64_t *nb = idx + i * ld_idx;
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
    for (int c0 = 0; c0 < d; c0 += 32 * DPL) {
      T acc[DPL];
#pragma unroll
      for (int q = 0; q < DPL; ++q) acc[q] = T(0);
      bool first = true;
#pragma unroll
      for (int kq = 0; kq < KPL; ++kq) {
        if (kq * 32 < k) {  // warp-uniform
          for (int jb = 0; jb < 32 && kq * 32 + jb < k; jb += SP_UNROLL) {
            // phase 1: all loads of the batch in flight (words of this sweep's rows, plain values of the others)
            T v[SP_UNROLL][DPL];
            unsigned long long w[SP_UNROLL][DPL][W];
            const unsigned long long *wsrc[SP_UNROLL];
#pragma unroll
            for (int u = 0; u < SP_UNROLL; ++u) {
              const int j = __shfl_sync(FULL, mine[kq], (jb + u) & 31);
              wsrc[u] = nullptr;
              if (kq * 32 + jb + u < k) {
                if (j < i32) {  // a row of THIS sweep: every word carries the epoch it was written in
                  wsrc[u] = next + static_cast<int64_t>(j) * d * W;
#pragma unroll
                  for (int q = 0; q < DPL; ++q) {
                    const int c = c0 + q * 32 + lane;
#pragma unroll
                    for (int e = 0; e < W; ++e) w[u][q][e] = (c < d) ? ld_word(wsrc[u] + c * W + e) : 0ull;
                  }
                } else {
                  const T *src = prev + static_cast<int64_t>(j) * d;
#pragma unroll
                  for (int q = 0; q < DPL; ++q) {
                    const int c = c0 + q * 32 + lane;
                    v[u][q] = (c < d) ? __ldg(src + c) : T(0);
                  }
                }
              }
            }
            // phase 2: a word without the current epoch is still in flight (here, or on its way over NVLink): re-read it
#pragma unroll
            for (int u = 0; u < SP_UNROLL; ++u) {
              if (wsrc[u] != nullptr) {
#pragma unroll
                for (int q = 0; q < DPL; ++q) {
                  const int c = c0 + q * 32 + lane;
                  T val = T(0);
                  if (c < d) {
                    unsigned ns = 32;
                    while (!LL<T>::ready(w[u][q], epoch)) {
                      __nanosleep(ns);
                      if (ns < 1024) ns <<= 1;
#pragma unroll
                      for (int e = 0; e < W; ++e) w[u][q][e] = ld_word(wsrc[u] + c * W + e);
                    }
                    val = LL<T>::value(w[u][q]);
                  }
                  v[u][q] = val;
                }
              }
            }
#pragma unroll
            for (int u = 0; u < SP_UNROLL; ++u) {
              if (kq * 32 + jb + u < k) {
#pragma unroll
                for (int q = 0; q < DPL; ++q) acc[q] = first ? v[u][q] : ArithP<T>::add(acc[q], v[u][q]);
                first = false;
              }
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < DPL; ++q) {
        const int c = c0 + q * 32 + lane;
        if (c < d) {
          const T r = __ldg(prev + i * d + c);
          const T target = ArithP<T>::div(acc[q], kk);
          const T val = ArithP<T>::add(r, ArithP<T>::mul(ArithP<T>::sub(target, r), coef));
          unsigned long long w[W];
          LL<T>::pack(val, epoch, w);
          for (int pr = 0; pr < world; ++pr)  // the local copy and the push to every peer: value and epoch in one store
#pragma unroll
            for (int e = 0; e < W; ++e) st_word(pp.ll[pr] + (i * d + c) * W + e, w[e]);
        }
      }
    }
    // Not needed for correctness (every word carries its own epoch), only for timeliness: without it the stores to
    // the peers linger in the write path and the rows waiting for them on the other GPUs wait ~100 us each
    // (8.4 ms per sweep); a gpu-scope fence drains them at a hundredth of the cost of a system-scope one.
    __threadfence();
  }
}

// {value, epoch} words -> plain values (after the barrier that follows a sweep); a word of another epoch raises *bad
template <typename T>
__global__ void __launch_bounds__(256) smooth_peer_unpack_kernel(const unsigned long long *__restrict__ ll, int64_t count,
                                                                unsigned epoch, T *__restrict__ out, int *bad) {
  constexpr int W = LL<T>::W;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    unsigned long long w[W];
#pragma unroll
    for (int e = 0; e < W; ++e) w[e] = ll[i * W + e];
    if (!LL<T>::ready(w, epoch)) *bad = 2;
    out[i] = LL<T>::value(w);
  }
}

struct PeerWs {
  int *bad;
  unsigned long long *ticket;
};

size_t smooth_peer_workspace_bytes() { return 512; }

template <typename T, int DPL>
static int launch_peer_k(const T *x, const PeerPtrs &pp, const int64_t *idx, int64_t ld_idx, int64_t n, int d, int k,
                         double coef, unsigned epoch, int rank, int world, PeerWs w, cudaStream_t st) {
  auto kern2 = smooth_peer_kernel<T, DPL, 2>;
  auto kern4 = smooth_peer_kernel<T, DPL, 4>;
  auto kern = k <= 64 ? kern2 : kern4;
  int bps = 0;
  VVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, SP_THREADS, 0));
  if (bps < 1) bps = 1;
  const int64_t rows = (n - rank + world - 1) / world;
  int64_t grid = static_cast<int64_t>(sm_count()) * bps;
  const int64_t useful = (rows + SP_THREADS / 32 - 1) / (SP_THREADS / 32);
  if (grid > useful) grid = useful;
  // rows in flight wait on each other (now also across GPUs): keep the window a small fraction of n
  const int64_t cap = n / (16 * (SP_THREADS / 32) * world) + 1;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  kern<<<static_cast<unsigned>(grid), SP_THREADS, 0, st>>>(x, pp, idx, ld_idx, n, d, k, static_cast<T>(coef), epoch, rank,
                                                          world, w.ticket, w.bad);
  VVB_CHECK_LAUNCH();
  return VVB_OK;
}

// One sweep.  x: this rank's copy of the sweep's INPUT (n, d); ll_ptrs: `world` device pointers, every rank's output
// copy as mapped into THIS process (entry `rank` is local): n * d words of 8 bytes for float32, twice that for float64,
// zero at allocation; `epoch` >= 1 and larger than in any earlier sweep on these buffers (so nothing is ever cleared).
// The caller brackets the launch with a barrier over the ranks on both sides (nobody still reads the buffers of the
// previous sweep before any rank pushes; all pushes landed before any rank unpacks its copy).
// workspace: smooth_peer_workspace_bytes(), zeroed here.
int launch_smooth_peer(const void *x, int elem_size, int64_t n, int d, const int64_t *idx, int64_t ld_idx, int k,
                       double coef, int epoch, int rank, int world, void *const *ll_ptrs, void *workspace,
                       cudaStream_t st) {
  VVB_REQUIRE(elem_size == 4 || elem_size == 8, "smooth: elem_size must be 4 or 8");
  VVB_REQUIRE(n >= 1 && d >= 1 && k >= 1 && ld_idx >= k && epoch >= 1, "smooth (peer): bad shape arguments");
  VVB_REQUIRE(world >= 1 && world <= SP_MAX_RANKS && rank >= 0 && rank < world, "smooth (peer): bad rank / world size");
  VVB_SUPPORTED(k <= 128 && d <= 128 && n <= 0x7fffffffLL, "smooth (peer): k <= 128, d <= 128 only (k=%d, d=%d)", k, d);
  PeerPtrs pp;
  for (int r = 0; r < SP_MAX_RANKS; ++r) pp.ll[r] = r < world ? static_cast<unsigned long long *>(ll_ptrs[r]) : nullptr;
  PeerWs w;
  w.bad = static_cast<int *>(workspace);
  w.ticket = reinterpret_cast<unsigned long long *>(static_cast<char *>(workspace) + 256);
  VVB_CUDA(cudaMemsetAsync(workspace, 0, 512, st));
  const unsigned ep = static_cast<unsigned>(epoch);
#define VVB_SP(T)                                                                                                       \
  do {                                                                                                                  \
    const T *xp = static_cast<const T *>(x);                                                                            \
    if (d <= 32) return launch_peer_k<T, 1>(xp, pp, idx, ld_idx, n, d, k, coef, ep, rank, world, w, st);                \
    if (d <= 64) return launch_peer_k<T, 2>(xp, pp, idx, ld_idx, n, d, k, coef, ep, rank, world, w, st);                \
    return launch_peer_k<T, 4>(xp, pp, idx, ld_idx, n, d, k, coef, ep, rank, world, w, st);                             \
  } while (0)
  if (elem_size == 4) VVB_SP(float);
  VVB_SP(double);
#undef VVB_SP
}

// The rank's own copy -> plain (n, d) values; workspace: the one of the sweep (its first int reports a word that did
// not carry `epoch`: a row that never arrived).
int launch_smooth_peer_unpack(const void *ll_local, int elem_size, int64_t count, int epoch, void *out, void *workspace,
                              cudaStream_t st) {
  VVB_REQUIRE(elem_size == 4 || elem_size == 8, "smooth: elem_size must be 4 or 8");
  const unsigned grid = static_cast<unsigned>(std::min<int64_t>((count + 255) / 256, sm_count() * 16));
  const unsigned long long *ll = static_cast<const unsigned long long *>(ll_local);
  if (elem_size == 4)
    smooth_peer_unpack_kernel<float><<<grid, 256, 0, st>>>(ll, count, static_cast<unsigned>(epoch), static_cast<float *>(out),
                                                          static_cast<int *>(workspace));
  else
    smooth_peer_unpack_kernel<double><<<grid, 256, 0, st>>>(ll, count, static_cast<unsigned>(epoch),
                                                           static_cast<double *>(out), static_cast<int *>(workspace));
  VVB_CHECK_LAUNCH();
  return VVB_OK;
}

}  // namespace vvb
