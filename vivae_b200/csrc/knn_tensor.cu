// Exact kNN, tensor-core path:  screening GEMM (tcgen05 / TMEM / TMA)  ->  FP32 re-rank
// + certificate  ->  brute-force fallback for the rows the certificate rejects.
//
// Screening key.  Rows are mean-centred and scaled by a power of two so that the largest
// centred norm is <= 64, then split into two FP16 numbers x^ = hi + lo (22 significant
// bits).  For query i and database row j the tensor cores evaluate, in FP32,
//     key(i,j) = |y^_j|^2 - 2 ( xh.yh + xh.yl + xl.yh )        ( = |x^-y^|^2 - |x^|^2 + O(2^-22) )
// as ONE accumulation: the query operand holds [-2 xh | 1 1 1] and [-2 xl], the database
// operand [yh | nh nl nl2] (the squared norm as a three-term FP16 sum riding in the padded
// K slots) and [yl]; three descriptor pairs (hi,hi) (hi,lo) (lo,hi) are issued into the
// same TMEM accumulator, so the epilogue is a bare compare of the accumulator against the
// row's running threshold.
//
// Kernel shape (one CTA per SM, persistent over the database):
//   CTA = 256 query rows (two M=128 MMAs share every database tile), database tile N=128,
//   warp 0 lane 0 : TMA producer  (database tiles through a 4-stage ring, in the CTA's cell order; it ends the
//                   sweep once the remaining cells are provably too far -- knn_cells.cu)
//   warp 1, one elected lane : tcgen05.mma issuer.  For d + 3 <= 64 the query operand lives in TENSOR MEMORY
//                   (128 columns, written once with tcgen05.st) and is the MMA's A operand from there (an
//                   M=128,N=128,K=16 MMA with both operands in shared memory reads 8 KB per 64 cycles = the
//                   whole 128 B/clk shared-memory bandwidth of the SM; with A in TMEM it reads 4 KB).  The
//                   remaining 384 columns hold THREE accumulator slots used round-robin by the (tile, M block)
//                   sequence.
//   warp 2        : TMEM allocator
//   warps 4..19   : epilogue -- tcgen05.ld 32x32b, lane = query row; TWO threads own a row (one per 64-column
//                   half of a tile), each with its running threshold tau and list length in registers and
//                   its own candidate list (global memory, L2 resident); a 32-column chunk is screened with a
//                   min-tree + one compare, survivors are appended with one 8-byte store; a nearly full list is
//                   compacted by the warp with a bitwise threshold selection, tightening tau.
// Everything a row ever rejected has key >= its final tau, which is what the certificate uses.
//
// Exactness.  The re-rank kernel recomputes the contract's FP32 fmaf-chain distance for the m
// candidates, sorts by (d2, index) and emits the k best.  Row i is certified when
//     s^2 * D_k(i)  <  (tau_i + |x^_i|^2 - eps_i) (1 - gamma_d)
//     eps_i   = tc_eps_rel(ka) (|x^_i| + Ymax)^2      screening error, grows with the K atoms `ka`
//     gamma_d = 2 (d + 3) 2^-24                        rounding of a rejected row's fmaf chain
// (derivation: DESIGN.md section 4, "Certificate"), i.e. even after the worst-case screening error
// and the worst-case rounding of the contract's own chain no rejected row can beat the k-th
// neighbour.  A row that fails is REFINED: a second sweep of the same kernel with the row's
// threshold set to T_i = s^2 D_k(i) (1 + gamma_d) - |x^_i|^2 + eps_i collects every database row
// whose key is below T_i -- that set provably contains every row tied with or better than the k-th
// candidate -- and the re-rank kernel orders it by (d2, index).  A refine list that fills up is
// reduced IN the sweep to its k+1 best entries by the contract's own key (exact fmaf chain, index),
// which also tightens T_i; with k+1 exact duplicates of the query in hand, later duplicates with a
// larger index are dropped on sight.  The brute-force kernel remains only as the guard behind it.
// Results are therefore identical to the brute-force contract whatever the tensor cores round.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "knn_cells.cuh"

namespace vvb {

int launch_knn_brute_list(const float *x_dev, int64_t n, int d, int k, const int64_t *qrows_dev,
                          const int *count_dev, int64_t max_rows, int64_t *out_idx, float *out_dist,
                          int64_t out_row0, void *workspace, size_t workspace_bytes, cudaStream_t stream);
size_t brute_workspace_bytes(int64_t nq, int k);
int knn_brute_few_max_rows();
size_t knn_brute_few_workspace_bytes(int k);
int launch_knn_brute_few(const float *x_dev, int64_t n, int d, int k, const int64_t *qrows_dev, int count,
                         int64_t *out_idx, float *out_dist, int64_t out_row0, void *workspace, size_t workspace_bytes,
                         cudaStream_t stream);

// ------------------------------------------------------------------ compile-time shape
#ifndef VVB_TC_STAGES
#define VVB_TC_STAGES 4
#endif
constexpr int TC_M = 256;            // query rows per CTA
constexpr int TC_N = 128;            // database rows per tile
constexpr int TC_STAGES = VVB_TC_STAGES;  // database-tile ring (32 KB per stage)
constexpr int TC_THREADS = 640;      // 4 control warps + 16 epilogue warps
constexpr int TC_ATOM_BYTES = 128 * 128;  // 128 rows x 64 fp16 (one 128B-swizzle K atom)
constexpr int TC_CAP_E = 8;          // candidate list capacity = 32 * 8 = 256
constexpr int TC_CAP = 32 * TC_CAP_E;
constexpr int TC_MAX_KA = 64;        // K atoms (64 columns each) per operand half: d + 3 <= 4096
constexpr int TC_KSTAGES = 2;        // K-loop variant: (query + database) atom stages
constexpr float TC_PAD_NORM = 60000.0f;  // squared "norm" of padding rows: their keys stay above TC_PAD_CUT
constexpr float TC_PAD_CUT = 30000.0f;   // real keys are below 64^2 + 2 * 64 * 64
constexpr int TC_RING = 16;              // tile-sequence ring (producer runs at most stages + slots tiles ahead)
constexpr int TC_PRE_TILES = TC_CAP;     // threshold pre-pass: at most this many tiles (two half-tile minima each, one
                                         // float per minimum, kept in the row's first 2 KB list buffer)
constexpr int TC_PRE_MAX_KA = 2;         // ... only where the epilogue, not the MMA, bounds the sweep (d + 3 <= 128)
constexpr int TC_SEQ_END = -1;           // tile_seq markers
constexpr int TC_SEQ_PRE_END = -2;
// Screening-error margin relative to (|x^| + Ymax)^2.  Model: every K=16 tcgen05.mma step
// commits at most 2^-21 of the accumulator's magnitude bound (|x^| + Ymax)^2 -- 3 * 4 * ka steps per key -- and the
// operand representation (centring, hi/lo split, dropped lo.lo term, FP32 norms) at most 2^-18 in total:
//     eps_rel(ka) = 2^-18 + 6 ka 2^-20         ka = 1: 2^-16.7,   ka = 8: 2^-14.3,   ka = 32 (d = 2000): 2^-12.4
// The tensor cores' internal rounding is undocumented, so the model is pinned by measurement
// (tests/test_gpu_configs.py::test_screening_error_stays_8x_below_the_margin_at_every_width): on B200 the error
// against FP64 is 2^-21.5 at d = 50, 2^-18.2 at d = 500, 2^-16.2 at d = 2000 -- linear in ka (truncation adds up,
// about 2^-24.8 per step), 14x to 28x below the margin.
__host__ __device__ constexpr float tc_eps_rel(int ka) {
  return 1.0f / 262144.0f + 6.0f * static_cast<float>(ka) / 1048576.0f;
}
// First tier of the K-loop variant: the key drops 2 (xh.yl + xl.yh + xl.yl); |xl| <= 2^-11 |x^| row-wise, so the
// dropped terms are at most 2 (2 * 2^-11 + 2^-22) |x^||y^| <= 1.001 * 2^-11 (|x^| + Ymax)^2 (|x||y| <= (|x| + |y|)^2 / 4).
__host__ __device__ constexpr float tc_eps_single(int ka) { return 1.001f / 2048.0f + tc_eps_rel(ka); }
// Rounding of the contract's own FP32 chain for a REJECTED row (d subtractions squared + d fmaf roundings, first
// order (d + 2) 2^-24 of the distance): doubled for the second-order terms.
static inline float tc_gamma(int d) { return 2.0f * static_cast<float>(d + 3) / 16777216.0f; }
constexpr int TC_SINGLE_MIN_KA = 4;   // rows of at least 190 columns: the sweep is bound by the MMAs, not by the epilogue
constexpr int TC_REFINE_CAP = 2 * TC_CAP;  // refine pass: candidates a row may collect below its fixed threshold

constexpr size_t TC_SMEM_B = TC_STAGES * 2 * TC_ATOM_BYTES;           // [stage][atom]
constexpr size_t TC_SMEM_TAIL = 256 /*barriers*/ + 8 * CELL_MAX /*cell order*/ + 4 * TC_RING + 64 /*tau*/ +
                                5 * TC_M * 4 /*per-row hand-over between the two warps of a row*/;
constexpr size_t TC_SMEM_BYTES = TC_SMEM_B + 1024 /*align*/ + TC_SMEM_TAIL;
constexpr int TC_TMEM_A = 384;       // resident variant: first TMEM column of the query operand (2 x 64 columns)
// K-loop variant: every stage holds one K atom of everything: query hi/lo for both M blocks + database hi/lo
constexpr size_t TC_SMEM_K = TC_KSTAGES * 6 * TC_ATOM_BYTES;
constexpr size_t TC_SMEM_BYTES_K = TC_SMEM_K + 1024 + TC_SMEM_TAIL;
static inline int tc_atoms(int d) { return (d + 3 + 63) / 64; }

static inline int tc_candidates(int k) {
  const int slack = k / 4 > 16 ? k / 4 : 16;
  int m = k + 1 + slack;
  // The re-rank kernel works in groups of 32 candidates (a 200-byte gather, an fmaf chain and a slot of the
  // bitonic sort each), and the screening kernel hands over between m and min(round_up_32(m), m + 8) of them:
  // a list that spills a few entries into one more group pays for the whole group.  If m is just above a
  // group boundary, step 4 below it (k = 50: 67 -> 60, at most 64 candidates: two groups instead of three and
  // a sort of 64 instead of 128 keys).  Exactness does not depend on m -- a row whose k-th distance is not
  // below its admission threshold goes to the brute-force fallback whatever m is.
  const int g = m & ~31;
  if (m - g <= 4 && g - 4 - (k + 1) >= 8) m = g - 4;
  return m;
}
// List length that triggers a compaction.  Measured on B200 (1M x 50, k=50; screening kernel time):
// trigger 224 of capacity 256: 380 ms; trigger 96 (fresher thresholds, 3x the compactions): 445 ms;
// capacity 512 / trigger 480: 433 ms.  A compaction stalls the whole warp (32 rows) and, through the
// accumulator hand-off, the MMA issuer, so it is done as late as the capacity allows.
static inline int tc_trigger(int keep) {
  (void)keep;
  return TC_CAP - 32;
}

bool knn_tensor_supported(int64_t n, int d, int k) {
  return d >= 1 && tc_atoms(d) <= TC_MAX_KA && tc_candidates(k) <= TC_CAP - 64 && n >= 2 && n <= 0x7fffffffLL;
}

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand in tensor memory (128 lanes x 8 columns of packed fp16 pairs per K=16 step), B in shared memory
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// tcgen05.ld is asynchronous: tc_ld32_issue starts the load of 32 consecutive columns of this
// thread's TMEM lane, tc_ld32_wait makes them usable.  The wait lists the registers as in/out
// operands so that the compiler cannot move their uses above it.
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld32_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                 "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                 "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// K-major, 128B-swizzle shared-memory descriptor of one atom (8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fff);  // start address
  d |= static_cast<uint64_t>(0) << 16;                     // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;             // stride byte offset
  d |= static_cast<uint64_t>(1) << 46;                     // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                     // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: FP16 x FP16 -> FP32, both operands K-major, M=128
__host__ __device__ constexpr uint32_t make_idesc(int n_cols) {
  return (1u << 4)                                   // c_format = F32
         | (0u << 7) | (0u << 10)                    // a_format = b_format = F16
         | (0u << 15) | (0u << 16)                   // K-major A and B
         | (static_cast<uint32_t>(n_cols >> 3) << 17)  // N
         | (static_cast<uint32_t>(128 >> 4) << 24);    // M
}

// ------------------------------------------------------------------ preparation kernels
constexpr int PREP_BLOCKS = 256;

// stage 1 of the column means: per-block partial sums in double; blockIdx.y = 64-column block
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float *__restrict__ x, int64_t n, int d,
                                                            double *__restrict__ partial) {
  // thread (c = threadIdx.x % 64, sub = threadIdx.x / 64) sums rows sub, sub+4*gridDim.x ... of its column
  const int cl = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + cl;
  double acc = 0.0;
  if (c < d)
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * 4 + sub; r < n; r += static_cast<int64_t>(gridDim.x) * 4)
      acc += static_cast<double>(__ldg(x + r * d + c));
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  if (sub == 0)
    partial[(static_cast<size_t>(blockIdx.x) * gridDim.y + blockIdx.y) * 64 + cl] =
        sh[cl] + sh[64 + cl] + sh[128 + cl] + sh[192 + cl];
}
__global__ void colsum_final_kernel(const double *__restrict__ partial, int nblocks, int64_t n, int d,
                                    float *__restrict__ mu, unsigned int *__restrict__ maxn2_bits) {
  const int cl = threadIdx.x;  // 64 threads; blockIdx.x = 64-column block
  const int c = blockIdx.x * 64 + cl;
  double acc = 0.0;
  for (int b = 0; b < nblocks; ++b) acc += partial[(static_cast<size_t>(b) * gridDim.x + blockIdx.x) * 64 + cl];
  mu[c] = (c < d) ? static_cast<float>(acc / static_cast<double>(n)) : 0.0f;
  if (c == 0) *maxn2_bits = 0u;
}
// largest squared centred norm (order-independent: atomicMax on the bits of a non-negative float)
__global__ void __launch_bounds__(256) maxnorm_kernel(const float *__restrict__ x, int64_t n, int d,
                                                     const float *__restrict__ mu,
                                                     unsigned int *__restrict__ maxn2_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  float best = 0.0f;
  for (int64_t r = warp; r < n; r += nwarps) {
    float s = 0.0f;
    for (int c = lane; c < d; c += 32) {
      const float t = __ldg(x + r * d + c) - __ldg(mu + c);
      s = fmaf(t, t, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    best = fmaxf(best, s);
  }
  if (lane == 0 && best > 0.0f) atomicMax(maxn2_bits, __float_as_uint(best));
}

struct PrepScalars {  // written by scalars_kernel, read by the operand, bound and re-rank kernels
  float scale;        // power of two
  float ymax;         // largest scaled centred norm (<= 64 before rounding slack)
};

__device__ __forceinline__ float pick_scale(float maxn2) {
  if (!(maxn2 > 0.0f)) return 1.0f;
  const float target = 64.0f / sqrtf(maxn2);
  int e;
  (void)frexpf(target, &e);  // target = m * 2^e, m in [0.5, 1)
  return ldexpf(1.0f, e - 1);  // largest power of two <= target
}

__global__ void scalars_kernel(const unsigned int *__restrict__ maxn2_bits, PrepScalars *__restrict__ scalars) {
  const float maxn2 = __uint_as_float(*maxn2_bits);
  const float s = pick_scale(maxn2);
  scalars->scale = s;
  scalars->ymax = sqrtf(maxn2) * s * 1.0001f;
}

// One warp per operand row: row w of the operand is original row list[w] (-1 = padding): centre, scale,
// split into FP16 hi/lo.  IS_QUERY: the query role ([-2 xh | 1 1 1], [-2 xl], and |x^|^2 to xnorm);
// otherwise the database role ([yh | nh nl nl2], [yl]).
// Operand row layout (fp16): [hi part: ka atoms of 64 columns | lo part: ka atoms]; the hi part carries
// the data in columns 0..d-1 and, in columns d..d+2, the norm terms (database) or ones (query).
template <bool IS_QUERY>
__global__ void __launch_bounds__(256) build_operands_kernel(const float *__restrict__ x, int d, int ka,
                                                            const float *__restrict__ mu,
                                                            const PrepScalars *__restrict__ scalars,
                                                            const int *__restrict__ list, int64_t rows,
                                                            __half *__restrict__ op, float *__restrict__ xnorm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const float s = scalars->scale;
  const int half_cols = ka * 64;
  const int64_t op_cols = 2 * half_cols;
  for (int64_t w = warp; w < rows; w += nwarps) {
    const int64_t src = list[w];
    const bool valid = src >= 0;
    __half *row = op + w * op_cols;
    const float fac = IS_QUERY ? -2.0f : 1.0f;  // power of two: exact in FP16
    float nrm = 0.0f;
    for (int c = lane; c < half_cols; c += 32) {
      float v = 0.0f;
      if (valid && c < d) v = (__ldg(x + src * d + c) - __ldg(mu + c)) * s;
      const __half hi = __float2half_rn(v);
      const float hf = __half2float(hi);
      const __half lo = __float2half_rn(v - hf);
      const float rep = hf + __half2float(lo);
      nrm = fmaf(rep, rep, nrm);
      row[c] = __float2half_rn(fac * hf);
      row[half_cols + c] = __float2half_rn(fac * __half2float(lo));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(FULL, nrm, o);
    __syncwarp();
    if (IS_QUERY) {
      if (lane == 0) xnorm[w] = nrm;
      if (lane < 3) row[d + lane] = __float2half_rn(valid ? 1.0f : 0.0f);
    } else {
      // squared norm as a three-term FP16 sum (33 significant bits); padding rows get a huge norm
      const float base = valid ? nrm : TC_PAD_NORM;
      const __half n0 = __float2half_rn(base);
      const float r1 = base - __half2float(n0);
      const __half n1 = __float2half_rn(r1);
      const __half n2 = __float2half_rn(r1 - __half2float(n1));
      if (lane < 3) row[d + lane] = (lane == 0) ? n0 : (lane == 1) ? n1 : n2;
    }
  }
}

// Out-of-line list maintenance, so that the hot screening loop stays small in the instruction
// cache.  In-sweep compaction uses the cheap threshold selection; the full sort runs once per
// row at the end of the sweep.
__device__ __noinline__ void tc_select_threshold_rows(const float *base, int r0, int r1, int cnt, int m, int lane,
                                                      float *out) {
  warp_select_threshold_rows_q15<2 * TC_CAP_E>(base, 2 * TC_CAP * 2, r0, r1, cnt, m, 3, TC_PAD_CUT, lane, out);
}
__device__ __noinline__ float tc_select_truncate_wide(uint2 *list, int cnt, int keep, int keep_max, int lane,
                                                      int *new_cnt) {  // up to 2 * TC_CAP entries
  return warp_select_truncate_pairs<2 * TC_CAP_E>(list, cnt, keep, keep_max, lane, new_cnt);
}
__device__ __noinline__ float tc_select_truncate(uint2 *list, int cnt, int keep, int keep_max, int lane, int *new_cnt) {
  return warp_select_truncate_pairs<TC_CAP_E>(list, cnt, keep, keep_max, lane, new_cnt);
}
// end of the sweep: the row's two half lists (la: ca entries, lb: cb entries) are truncated straight from where they
// are -- no append pass, which was two dependent global round trips per row -- and the result lands in la
__device__ __noinline__ float tc_select_truncate_halves_short(uint2 *la, int ca, const uint2 *lb, int cb, int keep,
                                                              int keep_max, int lane, int *new_cnt) {  // <= 128 entries
  return warp_select_truncate_pairs_2<TC_CAP_E / 2>(la, ca, lb, cb, keep, keep_max, lane, new_cnt);
}
__device__ __noinline__ float tc_select_truncate_halves(uint2 *la, int ca, const uint2 *lb, int cb, int keep, int keep_max,
                                                        int lane, int *new_cnt) {  // <= 256 entries
  return warp_select_truncate_pairs_2<TC_CAP_E>(la, ca, lb, cb, keep, keep_max, lane, new_cnt);
}


// Refine pass: reduce a full list to its `keepn` best entries by the CONTRACT's key (fmaf-chain d2, original index;
// the query itself first).  Warp-cooperative, warp-uniform arguments.  Entries that were beaten by keepn others on
// the exact key can never be among the row's k nearest, so dropping them is exact.  The key field of the kept
// entries is overwritten with their exact d2 (nothing reads screening keys after a refine sweep).
// Returns the new length; *dstar / *jstar = exact d2 and original index of the worst kept entry.
__device__ __noinline__ int tc_refine_compact(uint2 *list, int cnt, int keepn, const float *__restrict__ x, int d,
                                              const int *__restrict__ perm, int64_t g, int lane, float *dstar,
                                              int *jstar) {
  const float *xq = x + g * d;
  __syncwarp();
  for (int pos = lane; pos < cnt; pos += 32) {  // pass 1: the contract's chain for every entry
    const int j = perm[list[pos].y];
    const float *y = x + static_cast<int64_t>(j) * d;
    float acc = 0.0f;
    for (int c = 0; c < d; ++c) {
      const float df = __fsub_rn(__ldg(xq + c), __ldg(y + c));
      acc = __fmaf_rn(df, df, acc);
    }
    list[pos].x = (static_cast<int64_t>(j) == g) ? 0u : __float_as_uint(acc) + 1u;  // d2 >= 0: bits are monotone
  }
  __syncwarp();
  uint64_t key[TC_CAP_E];
  uint2 ent[TC_CAP_E];
#pragma unroll
  for (int e = 0; e < TC_CAP_E; ++e) {
    const int pos = e * 32 + lane;
    key[e] = ~0ull;
    ent[e] = make_uint2(0u, 0u);
    if (pos < cnt) {
      ent[e] = list[pos];
      key[e] = (static_cast<uint64_t>(ent[e].x) << 32) | static_cast<uint32_t>(perm[ent[e].y]);
    }
  }
  // keepn-th smallest key (keys are distinct: distinct original rows), bitwise from the top
  uint64_t T = 0;  // invariant: count(key < T) < keepn
  for (int b = 63; b >= 0; --b) {
    const uint64_t trial = T | (1ull << b);
    int c = 0;
#pragma unroll
    for (int e = 0; e < TC_CAP_E; ++e) c += (key[e] < trial) ? 1 : 0;
    c = __reduce_add_sync(FULL, c);
    if (c < keepn) T = trial;
  }
  __syncwarp();
  int base = 0;
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int e = 0; e < TC_CAP_E; ++e) {
    const bool keep = key[e] <= T;  // padding keys are ~0 > T (cnt > keepn)
    const unsigned kb = __ballot_sync(FULL, keep);
    if (keep) list[base + __popc(kb & lt_mask)] = ent[e];
    base += __popc(kb);
  }
  __syncwarp();
  const uint32_t hi = static_cast<uint32_t>(T >> 32);
  *dstar = hi ? __uint_as_float(hi - 1u) : 0.0f;
  *jstar = static_cast<int>(static_cast<uint32_t>(T));
  return base;
}

// ------------------------------------------------------------------ the screening kernel
struct TcParams {
  int64_t nq;       // query rows of this launch (entries of qlist)
  int ks_hi;        // k-steps (of 16) covering d + 3 columns
  int ks_lo;        // k-steps covering d columns
  int keep;         // m: candidates kept per row
  int trigger;      // a list longer than this is compacted before the next 32-column chunk
  int ka;           // K atoms per operand half (1: query operand resident; >1: K-loop variant)
  int debug_mode;   // profiling only (env VVB_TC_DEBUG): 1 = drain TMEM but skip the screening,
                    // 2 = screening minima only, no admissions, 3 = no TMEM drain at all.  Results are garbage.
                    // 4 = stress test: pseudo-random delays in the producer and before every accumulator hand-over
                    //     (results must stay exact: tests/test_gpu_configs.py)
  int n_cells;      // P
  float eps_rel;    // screening-error margin relative to (|x^| + Ymax)^2: tc_eps_rel(ka), or tc_eps_single(ka)
  int single;       // K-loop variant, first tier: ONE product (xh . yh) instead of three -- a third of the MMAs and half
                    // the operand traffic for a key that is good to 2^-11 (|x^| + Ymax)^2; rows whose certificate
                    // fails at that margin are re-swept at full precision by the refine pass
  // Refine instantiation only: the threshold of query-list entry t starts at fixed_thr[qlist[t] - fixed_q0]; no
  // pre-pass; a list that fills up is reduced to its k1 best entries by the exact key (tc_refine_compact).
  const float *fixed_thr;
  int64_t fixed_q0;
  // A refine launch of only a few blocks is spread over the GPU by slicing the DATABASE: gridDim.y = slices, block
  // (x, y) sweeps every slices-th tile of block x's order; the slices append to the SAME per-row lists through global
  // counters slice_cnt[2 * t + half] (no compaction: a list that overflows sends its row to the brute-force kernel)
  // and refine_concat_kernel joins the two halves afterwards.
  int slices;
  int *slice_cnt;
  // WIDE instantiation only: every query block is swept by `wslices` CTAs, CTA (block, w) taking every wslices-th tile
  // of the order (blockIdx.x = block * wslices + w, so that the CTAs resident at the same time share query operands:
  // 148 blocks x 1 MB of query operand at d = 2000 do not fit L2 and every tile re-read came from DRAM; 37 blocks x 4
  // slices do).  Each CTA keeps lists of its own -- list row (block * wslices + w) * TC_M + row, i.e. cand / cand_cnt /
  // cand_thr are (nq_pad * wslices) long -- and wide_merge_kernel joins them afterwards.
  int wslices;
  // Resident variant: the threshold pre-pass evaluates ONE product (xh . yh, the norm riding along) instead of three
  // and loads only the hi atom of every tile -- a third of the MMAs for a key good to tc_eps_single(1) of
  // (|x^| + Ymax)^2; the starting threshold it yields is raised by that much (plus the full key's own margin), so it
  // still bounds the m-th full-precision key from above.  Exactness never depends on the starting threshold.
  int pre_hi;
  float pre_slack;   // what the one-product threshold is raised by, relative to (|x^| + Ymax)^2
  // Non-refine, non-WIDE launches: the join of a row's two half lists (final truncation to the m best) is left to
  // halves_merge_kernel, which runs at full occupancy, instead of 16 warps doing it row by row while the SM's tensor
  // pipe waits (86k of 1,046k clocks per block at 1M x 50).  The halves' lengths go to slice_cnt[2 t + half], their
  // final thresholds to half_tau[2 t + half].
  int defer_merge;
  float *half_tau;
  const float *x;    // (n, d) the original rows
  const int *perm;   // permuted database row -> original row
  int d, k1;         // k1 = k + 1 (self included)
  float gamma;       // tc_gamma(d)
  const uint32_t *op_a;  // query operand rows as 32-bit words (64 per row), resident variant only
  const int *cta_order;  // (gridDim.x) query block served by every CTA, longest estimated sweep first; or nullptr
  const int *qlist;      // (nq_pad) original row of every query-list entry, -1 = padding
  const int *cell_of;    // (n) cell of every original row
  const int *cell_off;   // (P + 1) first permuted database row of every cell (multiples of TC_N)
  const float *cell_lb;  // (P, P) lower bound of the scaled distance between rows of two cells
  const float *xnorm;    // (nq_pad) |x^|^2 of the query rows
  const PrepScalars *scalars;
  unsigned long long *tiles_done;  // database tiles actually swept, summed over CTAs
  long long *times;      // profiling only (env VVB_TC_TIMES): 8 clock64 stamps per CTA, or nullptr
  uint2 *cand;      // (nq_pad, TC_CAP) candidate lists: (key bits, PERMUTED database row id)
  int *cand_cnt;    // (nq_pad)
  float *cand_thr;  // (nq_pad) final admission threshold tau
};

// KLOOP = false: d + 3 <= 64.  The query operand (256 rows x [hi 64 | lo 64] fp16) is copied once into
//                 tensor memory (columns 384..511) and feeds the MMAs from there; only database tiles
//                 stream through shared memory (4-stage ring).  Three accumulator slots (columns 0..383).
// KLOOP = true : any d; every database tile is accumulated over `ka` K atoms, query and database atoms
//                stream together through a 2-stage ring (6 atoms = 96 KB per stage); four accumulator
//                slots.  The epilogue is identical.
// Accumulator slots: the sequence it = 2 t + mb (tile t, M block mb) uses slot it % NSLOT for the
// (it / NSLOT)-th time; acc_full[slot] is committed by the MMA issuer, acc_empty[slot] is released by
// the four epilogue warps of that M block.
//
// Sweep order.  The query rows of a CTA are consecutive entries of the cell-sorted query list, the
// database is stored cell by cell (knn_cells.cu).  The CTA sorts the cells by its lower bound of the
// distance to them and the producer walks that order, publishing the first database row of every tile
// it loads in `tile_seq` (-1 = end of sweep).  Before entering a cell with a positive bound it reads
// the largest "k-th candidate distance" the epilogue warps currently hold (tau_i + |x^_i|^2, refreshed
// at every list compaction): once bound^2 exceeds it (plus the screening error), no row of this or
// any later cell can enter any list and the sweep ends.
//
// Threshold pre-pass.  With the far cells skipped, what a CTA sweeps is its own neighbourhood, where
// admissions are dense: starting every row at tau = +inf costs ~900 list insertions and ~5 compactions
// per row before tau has decayed to the k-th neighbour distance, and that list handling, not the MMA,
// then bounds the kernel.  So the first TC_PRE_TILES tiles of the order are swept TWICE: the first time
// the epilogue only records each tile's minimum key per row (four 3-input-min trees per tile, no
// lists).  The m-th smallest of those minima is a valid starting threshold -- m different tiles each
// hold a key at or below it -- and a tight one, because the m nearest rows of a query are spread over
// almost as many tiles.  The real sweep then starts at that threshold: ~100 insertions per row, rarely a
// compaction.  Exactness never depends on it (a threshold is only ever an upper bound of the m-th key).
// WIDE (K-loop variant, first tier only): database tiles of 256 rows.  The K-loop re-streams the block's whole query
// operand (d x 256 rows) for every database tile, and at d = 2000 that stream -- not the MMAs -- bounds the sweep
// (148 blocks x 1 MB exceed L2, so it comes from DRAM); a 256-row tile halves it.  One M=128, N=256 MMA per M block
// and k-step, accumulators = all 512 TMEM columns (columns [256 mb, 256 mb + 256)), no double buffering: a tile is
// ka x 1024 MMA clocks, its read-out under 1000.  Stage = [A.hi mb0][A.hi mb1][B.hi rows 0..127][B.hi rows 128..255]
// (64 KB, three stages).  The last tile of a cell may hold 128 rows only (bit 0 of its tile_seq entry).
template <bool KLOOP, bool DBG, bool REFINE, bool WIDE = false>
__global__ void __launch_bounds__(TC_THREADS, 1) knn_screen_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                   const __grid_constant__ CUtensorMap map_b,
                                                                   const TcParams p) {
  constexpr int NSLOT = KLOOP ? 4 : 3;
  const int dbg_mode = DBG ? p.debug_mode : 0;  // the profiling modes live in their own instantiation
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // 128B swizzle needs 1024-byte aligned atoms
  const uint32_t smem_b = base;                                 // [stage][atom 2] x 16 KB  (resident variant)
  const uint32_t smem_k = base;                                 // [stage][6 atoms]         (K-loop variant)
  const uint32_t bars = base + (KLOOP ? TC_SMEM_K : TC_SMEM_B);  // mbarriers (8 B each)
  const uint32_t bar_a_full = bars;
  const uint32_t bar_b_full = bars + 8;                         // [TC_STAGES]
  const uint32_t bar_b_empty = bar_b_full + 8 * TC_STAGES;      // [TC_STAGES]
  const uint32_t bar_acc_full = bar_b_empty + 8 * TC_STAGES;    // [4]
  const uint32_t bar_acc_empty = bar_acc_full + 32;             // [4]
  const uint32_t tmem_slot = bar_acc_empty + 32;                // u32 written by tcgen05.alloc
  uint8_t *tail = smem_raw + (bars - raw) + 256;
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - raw));
  unsigned long long *order = reinterpret_cast<unsigned long long *>(tail);             // [CELL_MAX] (bound, cell)
  volatile int *tile_seq = reinterpret_cast<volatile int *>(tail + 8 * CELL_MAX);       // [TC_RING]
  volatile float *tau_pub = reinterpret_cast<volatile float *>(tail + 8 * CELL_MAX + 4 * TC_RING);  // [16]
  float *row_tau = reinterpret_cast<float *>(tail + 8 * CELL_MAX + 4 * TC_RING + 64);  // [TC_M] pre-pass threshold
  int *half_cnt = reinterpret_cast<int *>(row_tau + TC_M);                            // [2][TC_M] final list lengths
  float *half_tau = reinterpret_cast<float *>(half_cnt + 2 * TC_M);                   // [2][TC_M] final thresholds

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wsl = WIDE ? p.wslices : 1;                      // CTAs per query block (database slices)
  const int w_slice = static_cast<int>(blockIdx.x) % wsl;   // this CTA's slice
  const int q_block = p.cta_order ? p.cta_order[blockIdx.x / wsl] : static_cast<int>(blockIdx.x) / wsl;
  const int64_t q_base = static_cast<int64_t>(q_block) * TC_M;  // first query-list entry of this CTA
  // list row of query-list entry t: t + l_off (one set of lists per slice)
  const int64_t l_off = (static_cast<int64_t>(q_block) * (wsl - 1) + w_slice) * TC_M;
  auto stamp = [&](int i) {
    if (DBG && p.times && warp == 4 && lane == 0) p.times[q_block * 8 + i] = clock64();
  };
  stamp(0);

  if (warp == 0 && lane == 0) {
    if constexpr (KLOOP) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar_a_full, 8);  // resident variant: one arrival per epilogue warp once its rows are in TMEM
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_b_full + 8 * s, 1);
      mbar_init(bar_b_empty + 8 * s, 1);
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(bar_acc_full + 8 * b, 1);
      mbar_init(bar_acc_empty + 8 * b, 8);  // one arrival per epilogue warp of the M block
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < 16) tau_pub[threadIdx.x] = __int_as_float(0x7f800000);

  // ---- cell order of this CTA: bound(b) = min over the cells a of its rows of lb(a, b), sorted ascending
  const int P = p.n_cells;
  int P2 = 1;
  while (P2 < P) P2 <<= 1;
  {
    const int64_t last = min(q_base + TC_M, p.nq) - 1;
    const int a0 = p.cell_of[p.qlist[q_base]], a1 = p.cell_of[p.qlist[last]];  // the list is sorted by cell
    for (int b = threadIdx.x; b < P2; b += TC_THREADS) {
      float v = __int_as_float(0x7f800000);
      if (b < P)
        for (int a = a0; a <= a1; ++a) v = fminf(v, __ldg(p.cell_lb + static_cast<size_t>(a) * P + b));
      // WIDE: cells that cannot be skipped anyway (bound <= 0) are swept in cell order by EVERY block, so that the
      // blocks resident at the same time read the same database tiles and all but the first hit L2
      if (WIDE && v <= 0.0f) v = 0.0f;
      order[b] = (static_cast<unsigned long long>(f2ord(v)) << 32) | static_cast<unsigned>(b);
    }
  }
  __syncthreads();
  for (int size = 2; size <= P2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (P2 >> 1); i += TC_THREADS) {
        const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
        const int hi = lo | stride;
        const bool up = (lo & size) == 0;
        const unsigned long long x0 = order[lo], x1 = order[hi];
        if ((x0 > x1) == up) {
          order[lo] = x1;
          order[hi] = x0;
        }
      }
      __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  stamp(1);
  const bool pre_pass = (dbg_mode == 0 || dbg_mode >= 4) && p.ka <= TC_PRE_MAX_KA && !REFINE;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      const float ymax2 = 2.0f * p.scalars->ymax;
      const float eps_max = p.eps_rel * ymax2 * ymax2;  // the certificate's margin for the longest query row
      int t = 0;    // tile_seq entries issued (tiles and markers)
      int it = 0;   // K-loop variant: stage uses
      int swept = 0, pre_swept = 0;
      int tile_no = 0;  // sliced refine launch: position of the tile in the block's order
      // a marker: a stage carrying no data, flagged in the tile sequence
      auto issue_marker = [&](int marker) {
        const int u = KLOOP ? it : t;
        const int ns = WIDE ? 3 : KLOOP ? (p.single ? 2 * TC_KSTAGES : TC_KSTAGES) : TC_STAGES;
        const int s = u % ns;
        const uint32_t ph = (u / ns) & 1;
        mbar_wait(bar_b_empty + 8 * s, ph ^ 1);
        tile_seq[t % TC_RING] = marker;
        mbar_arrive(bar_b_full + 8 * s);
        ++t;
        ++it;
      };
      // the pre-pass covers the block's own neighbourhood: the cells its bound cannot separate from it
      // (bound <= 0), at most TC_PRE_TILES tiles; with too few tiles for m distinct minima it is skipped
      int pre_tiles = 0;
      if (pre_pass) {
        for (int ci = 0; ci < P && pre_tiles < TC_PRE_TILES; ++ci) {
          const unsigned long long e = order[ci];
          if (ord2f(static_cast<uint32_t>(e >> 32)) > 0.0f) break;
          const int b = static_cast<int>(static_cast<uint32_t>(e));
          pre_tiles += (p.cell_off[b + 1] - p.cell_off[b]) / TC_N;
        }
        pre_tiles = min(pre_tiles, TC_PRE_TILES);
        if (pre_tiles < p.keep + 16) pre_tiles = 0;
      }
      for (int pass = pre_pass ? 0 : 1; pass < 2; ++pass) {
      int pass_tiles = 0;
      for (int ci = 0; ci < P; ++ci) {
        const unsigned long long e = order[ci];
        const float bound = ord2f(static_cast<uint32_t>(e >> 32));
        const int b = static_cast<int>(static_cast<uint32_t>(e));
        if (bound == __int_as_float(0x7f800000)) break;  // empty cells sort last
        if (pass == 0 && pass_tiles >= pre_tiles) break;
        if (pass == 1 && bound > 0.0f) {
          float tmax = tau_pub[0];
#pragma unroll
          for (int w = 1; w < 16; ++w) tmax = fmaxf(tmax, tau_pub[w]);
          if (bound * bound >= tmax + eps_max) break;  // ascending order: every later cell is at least as far
        }
        const int row_end = p.cell_off[b + 1];
        if constexpr (WIDE) {
          for (int row_b = p.cell_off[b]; row_b < row_end; row_b += 2 * TC_N, ++t) {
            if (wsl > 1 && (tile_no++ % wsl) != w_slice) {  // another slice's tile
              --t;
              continue;
            }
            const bool half_tile = row_b + TC_N >= row_end;
            pass_tiles += half_tile ? 1 : 2;
            for (int a = 0; a < p.ka; ++a, ++it) {
              const int s = it % 3;
              const uint32_t ph = (it / 3) & 1;
              const uint32_t st = smem_k + s * 4 * TC_ATOM_BYTES;
              mbar_wait(bar_b_empty + 8 * s, ph ^ 1);
              if (a == 0) tile_seq[t % TC_RING] = row_b | (half_tile ? 1 : 0);
              mbar_expect_tx(bar_b_full + 8 * s, (half_tile ? 3 : 4) * TC_ATOM_BYTES);
              tma_load_2d(st + 0 * TC_ATOM_BYTES, &map_a, bar_b_full + 8 * s, a * 64, static_cast<int>(q_base));
              tma_load_2d(st + 1 * TC_ATOM_BYTES, &map_a, bar_b_full + 8 * s, a * 64, static_cast<int>(q_base + 128));
              tma_load_2d(st + 2 * TC_ATOM_BYTES, &map_b, bar_b_full + 8 * s, a * 64, row_b);
              if (!half_tile) tma_load_2d(st + 3 * TC_ATOM_BYTES, &map_b, bar_b_full + 8 * s, a * 64, row_b + TC_N);
            }
          }
        }
        if constexpr (!WIDE)
        for (int row_b = p.cell_off[b]; row_b < row_end; row_b += TC_N, ++t, ++pass_tiles) {
          if (pass == 0 && pass_tiles >= pre_tiles) break;
          if constexpr (REFINE) {
            if (p.slices > 1 && (tile_no++ % p.slices) != static_cast<int>(blockIdx.y)) {  // another slice's tile
              --t;
              --pass_tiles;
              continue;
            }
          }
          if (DBG && dbg_mode == 4) __nanosleep((static_cast<unsigned>(t) * 2246822519u >> 21) & 0x7ffu);
          if constexpr (!KLOOP) {
            const int s = t % TC_STAGES;
            const uint32_t ph = (t / TC_STAGES) & 1;
            mbar_wait(bar_b_empty + 8 * s, ph ^ 1);
            tile_seq[t % TC_RING] = row_b;
            const int natoms = (pass == 0 && p.pre_hi) ? 1 : 2;  // pre-pass with one product: the hi atom only
            if (DBG && dbg_mode == 7 && pass == 0) {  // profiling only: no loads in the pre-pass (MMA rate alone)
              mbar_arrive(bar_b_full + 8 * s);
              continue;
            }
            mbar_expect_tx(bar_b_full + 8 * s, natoms * TC_ATOM_BYTES);
            for (int atom = 0; atom < natoms; ++atom)
              tma_load_2d(smem_b + (s * 2 + atom) * TC_ATOM_BYTES, &map_b, bar_b_full + 8 * s, atom * 64, row_b);
          } else {
            // first tier (p.single): a stage holds [A.hi mb0][A.hi mb1][B.hi] (48 KB, four stages);
            // full precision: [A.hi mb0][A.lo mb0][A.hi mb1][A.lo mb1][B.hi][B.lo] (96 KB, two stages)
            const int kst = p.single ? 2 * TC_KSTAGES : TC_KSTAGES;
            const int stage_atoms = p.single ? 3 : 6;
            for (int a = 0; a < p.ka; ++a, ++it) {
              const int s = it % kst;
              const uint32_t ph = (it / kst) & 1;
              const uint32_t st = smem_k + s * stage_atoms * TC_ATOM_BYTES;
              const int c_hi = a * 64, c_lo = (p.ka + a) * 64;
              mbar_wait(bar_b_empty + 8 * s, ph ^ 1);
              if (a == 0) tile_seq[t % TC_RING] = row_b;
              mbar_expect_tx(bar_b_full + 8 * s, stage_atoms * TC_ATOM_BYTES);
              if (p.single) {
                tma_load_2d(st + 0 * TC_ATOM_BYTES, &map_a, bar_b_full + 8 * s, c_hi, static_cast<int>(q_base));
                tma_load_2d(st + 1 * TC_ATOM_BYTES, &map_a, bar_b_full + 8 * s, c_hi, static_cast<int>(q_base + 128));
                tma_load_2d(st + 2 * TC_ATOM_BYTES, &map_b, bar_b_full + 8 * s, c_hi, row_b);
              } else {
                tma_load_2d(st + 0 * TC_ATOM_BYTES, &map_a, bar_b_full + 8 * s, c_hi, static_cast<int>(q_base));
                tma_load_2d(st + 1 * TC_ATOM_BYTES, &map_a, bar_b_full + 8 * s, c_lo, static_cast<int>(q_base));
                tma_load_2d(st + 2 * TC_ATOM_BYTES, &map_a, bar_b_full + 8 * s, c_hi, static_cast<int>(q_base + 128));
                tma_load_2d(st + 3 * TC_ATOM_BYTES, &map_a, bar_b_full + 8 * s, c_lo, static_cast<int>(q_base + 128));
                tma_load_2d(st + 4 * TC_ATOM_BYTES, &map_b, bar_b_full + 8 * s, c_hi, row_b);
                tma_load_2d(st + 5 * TC_ATOM_BYTES, &map_b, bar_b_full + 8 * s, c_lo, row_b);
              }
            }
          }
        }
      }
      swept += pass_tiles;
      if (pass == 0 && !KLOOP && p.pre_hi) pre_swept = pass_tiles;
      issue_marker(pass == 0 ? TC_SEQ_PRE_END : TC_SEQ_END);
      }
      atomicAdd(p.tiles_done, static_cast<unsigned long long>(swept));
      if (pre_swept) atomicAdd(p.tiles_done + 1, static_cast<unsigned long long>(pre_swept));  // one-product tiles
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // One ELECTED lane runs the loop (elect.sync, not a lane-id test: ptxas then keeps descriptors and
    // addresses on the uniform datapath and emits back-to-back UTCHMMAs; with `lane == 0` it wrapped
    // every MMA in a per-thread serialisation loop that cost ~100 cycles per MMA and bounded the kernel).
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(TC_N);
      if constexpr (!KLOOP) {
        mbar_wait(bar_a_full, 0);  // query operand is in tensor memory
        tc_fence_after();
      }
      const uint64_t desc_b0 = make_desc(KLOOP ? smem_k : smem_b);  // + (byte offset >> 4) moves the start address
      int slot = 0;
      uint32_t use_par = 0;  // parity of (it / NSLOT)
      int it = 0;            // K-loop variant: stage uses
      bool hi_only = !KLOOP && pre_pass && p.pre_hi;  // tiles before the TC_SEQ_PRE_END marker: one product
      if constexpr (WIDE) {
        constexpr uint32_t idesc2 = make_idesc(2 * TC_N);
        for (int t = 0;; ++t) {
          const uint32_t tpar = static_cast<uint32_t>(t) & 1u;
          mbar_wait(bar_b_full + 8 * (it % 3), (it / 3) & 1);
          const int marker = tile_seq[t % TC_RING];
          mbar_wait(bar_acc_empty + 0, tpar ^ 1);
          mbar_wait(bar_acc_empty + 8, tpar ^ 1);
          if (marker < 0) {  // end of the sweep: pass the marker on to both M blocks
            mbar_arrive(bar_acc_full + 0);
            mbar_arrive(bar_acc_full + 8);
            break;
          }
          for (int a = 0; a < p.ka; ++a, ++it) {
            const int s = it % 3;
            if (a > 0) mbar_wait(bar_b_full + 8 * s, (it / 3) & 1);
            tc_fence_after();
            const uint64_t st = desc_b0 + static_cast<uint64_t>(s * (4 * TC_ATOM_BYTES >> 4));
            const uint64_t b_hi = st + 2 * (TC_ATOM_BYTES >> 4);  // 256 rows: two atoms back to back
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
              const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(mb * 2 * TC_N);
              const uint64_t a_hi = st + mb * (TC_ATOM_BYTES >> 4);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                tc_mma_f16(d_tmem, a_hi + 2 * ks, b_hi + 2 * ks, idesc2, (a > 0 || ks > 0) ? 1u : 0u);
            }
            tc_commit(bar_b_empty + 8 * s);
          }
          tc_commit(bar_acc_full + 0);
          tc_commit(bar_acc_full + 8);
        }
      } else
      for (int t = 0;; ++t) {
        const int kst = p.single ? 2 * TC_KSTAGES : TC_KSTAGES;  // K-loop variant: stages in the ring
        const int s0 = KLOOP ? it % kst : t % TC_STAGES;
        const uint32_t ph0 = KLOOP ? (it / kst) & 1 : (t / TC_STAGES) & 1;
        mbar_wait(bar_b_full + 8 * s0, ph0);
        const int marker = tile_seq[t % TC_RING];
        if (marker < 0) {
          // pass the marker on (the epilogue reads tile_seq after its acc_full wait) and free the stage
          for (int mb = 0; mb < 2; ++mb) {
            mbar_wait(bar_acc_empty + 8 * slot, use_par ^ 1);
            mbar_arrive(bar_acc_full + 8 * slot);
            if (++slot == NSLOT) {
              slot = 0;
              use_par ^= 1;
            }
          }
          if (marker == TC_SEQ_END) break;
          hi_only = false;
          mbar_arrive(bar_b_empty + 8 * s0);
          ++it;
          continue;
        }
        if constexpr (!KLOOP) {
          const uint64_t b_hi = desc_b0 + static_cast<uint64_t>(s0 * (2 * TC_ATOM_BYTES >> 4));
          const uint64_t b_lo = b_hi + (TC_ATOM_BYTES >> 4);
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) {
            mbar_wait(bar_acc_empty + 8 * slot, use_par ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(slot * TC_N);
            const uint32_t a_hi = tmem_base + static_cast<uint32_t>(TC_TMEM_A + mb * 64);  // 8 columns per k-step
            const uint32_t a_lo = a_hi + 32;
            // [-2xh | 1 1 1] . [yh | norm],  xh . yl,  xl . yh   (k-steps beyond the data are skipped)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              if (ks < p.ks_hi) tc_mma_f16_ts(d_tmem, a_hi + 8 * ks, b_hi + 2 * ks, idesc, ks > 0 ? 1u : 0u);
            if (!hi_only) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                if (ks < p.ks_lo) tc_mma_f16_ts(d_tmem, a_hi + 8 * ks, b_lo + 2 * ks, idesc, 1u);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                if (ks < p.ks_lo) tc_mma_f16_ts(d_tmem, a_lo + 8 * ks, b_hi + 2 * ks, idesc, 1u);
            }
            tc_commit(bar_acc_full + 8 * slot);  // accumulator ready for the epilogue
            if (++slot == NSLOT) {
              slot = 0;
              use_par ^= 1;
            }
          }
          tc_commit(bar_b_empty + 8 * s0);  // smem stage reusable once these MMAs retire
        } else {
          // slots of this tile: slot (mb 0) and slot + 1 (mb 1); NSLOT = 4 keeps the pair aligned
          mbar_wait(bar_acc_empty + 8 * slot, use_par ^ 1);
          mbar_wait(bar_acc_empty + 8 * (slot + 1), use_par ^ 1);
          const int stage_atoms = p.single ? 3 : 6;
          for (int a = 0; a < p.ka; ++a, ++it) {
            const int s = it % kst;
            const uint32_t ph = (it / kst) & 1;
            if (a > 0) mbar_wait(bar_b_full + 8 * s, ph);
            tc_fence_after();
            const uint64_t st = desc_b0 + static_cast<uint64_t>(s * (stage_atoms * TC_ATOM_BYTES >> 4));
            if (p.single) {
              const uint64_t b_hi = st + 2 * (TC_ATOM_BYTES >> 4);
#pragma unroll
              for (int mb = 0; mb < 2; ++mb) {
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>((slot + mb) * TC_N);
                const uint64_t a_hi = st + mb * (TC_ATOM_BYTES >> 4);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  tc_mma_f16(d_tmem, a_hi + 2 * ks, b_hi + 2 * ks, idesc, (a > 0 || ks > 0) ? 1u : 0u);
              }
            } else {
              const uint64_t b_hi = st + 4 * (TC_ATOM_BYTES >> 4);
              const uint64_t b_lo = st + 5 * (TC_ATOM_BYTES >> 4);
#pragma unroll
              for (int mb = 0; mb < 2; ++mb) {
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>((slot + mb) * TC_N);
                const uint64_t a_hi = st + (mb * 2 + 0) * (TC_ATOM_BYTES >> 4);
                const uint64_t a_lo = st + (mb * 2 + 1) * (TC_ATOM_BYTES >> 4);
                // zero-padded columns contribute exactly 0, so every atom runs its 4 k-steps
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  tc_mma_f16(d_tmem, a_hi + 2 * ks, b_hi + 2 * ks, idesc, (a > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) tc_mma_f16(d_tmem, a_hi + 2 * ks, b_lo + 2 * ks, idesc, 1u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) tc_mma_f16(d_tmem, a_lo + 2 * ks, b_hi + 2 * ks, idesc, 1u);
              }
            }
            tc_commit(bar_b_empty + 8 * s);
          }
          tc_commit(bar_acc_full + 8 * slot);
          tc_commit(bar_acc_full + 8 * (slot + 1));
          slot += 2;
          if (slot == NSLOT) {
            slot = 0;
            use_par ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue: fused per-row top-m ===========================
    // 16 warps: a query row is owned by TWO threads (same lane of two warps of the same TMEM quadrant), one per
    // 64-column half of every database tile, each with its own candidate list, list length and threshold
    // (four epilogue warps per scheduler hide the dependent-issue latency of the list handling better than
    // two).  The halves meet twice: at the end of the pre-pass (one common threshold per row) and at the end of
    // the sweep (the second list is appended to the first); each warp of a pair serves 16 of the 32 rows there.
    const int ew = warp - 4;           // 0..15
    const int quad = warp & 3;         // TMEM lane quadrant this warp may access
    const int mb = (ew >> 2) & 1;      // which M=128 block
    const int half = ew >> 3;          // which 64 columns of every tile
    const int prow = mb * 128 + quad * 32 + lane;  // row within the CTA
    const int64_t t_row = q_base + prow;           // query-list entry owned by this thread
    const bool active = t_row < p.nq;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int pair_bar = 1 + (ew & 7);  // named barrier of the two warps that own the same 32 rows
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory"); };
    if constexpr (!KLOOP) {
      if (half == 0) {
        // this thread's query operand row -> tensor memory: 64 words = [hi: 32 columns | lo: 32 columns],
        // two fp16 per 32-bit column (K index 2c in the low half), which is the A-operand layout of kind::f16
        const uint4 *src = reinterpret_cast<const uint4 *>(p.op_a + t_row * 64);
        uint32_t w[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const uint4 q4 = __ldg(src + h * 8 + v);
            w[4 * v + 0] = q4.x;
            w[4 * v + 1] = q4.y;
            w[4 * v + 2] = q4.z;
            w[4 * v + 3] = q4.w;
          }
          tc_st32(t_lane + static_cast<uint32_t>(TC_TMEM_A + mb * 64 + h * 32), w);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a_full);
      }
    }
    uint2 *cl = p.cand + ((t_row + l_off) * 2 + half) * TC_CAP;  // this thread's candidate list
    // opaque to the optimiser: kept in a register pair, so that an admission is IMAD.WIDE + STG + count (the compiler
    // otherwise rebuilds the address from p.cand and the row index with four 64-bit steps per admission)
    asm volatile("" : "+l"(cl));
    uint2 *warp_lists = p.cand + ((q_base + l_off + mb * 128 + quad * 32) * 2 + half) * TC_CAP;  // + 2 * TC_CAP per lane
    const float inf = __int_as_float(0x7f800000);
    float tau = active ? inf : -inf;
    const float xn = active ? p.xnorm[t_row] : 0.0f;
    int cnt = 0;
    // refine state: once a compaction has found k1 exact duplicates of the query (dstar == 0), a later candidate
    // can only displace one of them with a smaller original index
    int dup_jmax = 0x7fffffff;
    float eps_row = 0.0f, scale2 = 1.0f;
    if constexpr (REFINE) {
      if (active) tau = p.fixed_thr[p.qlist[t_row] - p.fixed_q0];
      const float s = p.scalars->scale, rr = sqrtf(xn) + p.scalars->ymax;
      scale2 = s * s;
      eps_row = p.eps_rel * rr * rr * 1.0625f;
    }
    {  // a warp without query rows must not hold the sweep open
      const unsigned m0 = __reduce_max_sync(FULL, f2ord(REFINE ? tau + xn : tau));
      if (lane == 0) tau_pub[ew] = ord2f(m0);
    }

    // One 32-column chunk: make room, then screen.
    auto process = [&](const uint32_t(&r)[32], int j0) {
      if (dbg_mode == 1) {
        if (__uint_as_float(r[0]) == 123.456f && __uint_as_float(r[31]) == 1.5f) tau = 0.f;  // keep the load alive
        return;
      }
      unsigned need = __ballot_sync(FULL, cnt > p.trigger);
      if (REFINE && need) {
        // a full refine list keeps its k1 best entries by the exact key; the k1-th of them bounds the row's k-th
        // distance from above, which tightens the threshold (same formula as the re-rank kernel's fb_thr)
        while (need) {
          const int src = __ffs(need) - 1;
          need &= need - 1;
          const int c = __shfl_sync(FULL, cnt, src);
          const int64_t tr = __shfl_sync(FULL, t_row, src);
          float dstar;
          int jstar;
          const int nc = tc_refine_compact(warp_lists + src * 2 * TC_CAP, c, p.k1, p.x, p.d, p.perm, p.qlist[tr], lane,
                                           &dstar, &jstar);
          if (lane == src) {
            cnt = nc;
            float T = fmaf(dstar * scale2, 1.0f + p.gamma, eps_row - xn);
            T += fabsf(T) * 1e-6f + 1e-30f;
            tau = fminf(tau, T);
            dup_jmax = (dstar == 0.0f) ? jstar : 0x7fffffff;
          }
        }
        const unsigned m = __reduce_max_sync(FULL, f2ord(tau + xn));
        if (lane == 0) tau_pub[ew] = ord2f(m);
      } else if (need) {
        while (need) {
          const int src = __ffs(need) - 1;
          need &= need - 1;
          const int c = __shfl_sync(FULL, cnt, src);
          int nc;
          const float thr = tc_select_truncate(warp_lists + src * 2 * TC_CAP, c, p.keep, p.trigger, lane, &nc);
          if (lane == src) {
            cnt = nc;
            tau = thr;
          }
        }
        // publish the largest k-th candidate distance of the warp's rows (the producer's stopping rule)
        const unsigned m = __reduce_max_sync(FULL, f2ord(tau + xn));  // inactive rows: -inf
        if (lane == 0) tau_pub[ew] = ord2f(m);
      }
      // minima of 3 + 3 + 2 columns, of the groups of 8 and of the chunk; a hit is located by descending that tree
      const float cut = fminf(tau, TC_PAD_CUT);  // padding rows of the database carry a huge norm
      float g[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float a = fminf(fminf(__uint_as_float(r[8 * q + 0]), __uint_as_float(r[8 * q + 1])), __uint_as_float(r[8 * q + 2]));
        const float b = fminf(fminf(__uint_as_float(r[8 * q + 3]), __uint_as_float(r[8 * q + 4])), __uint_as_float(r[8 * q + 5]));
        g[q] = fminf(fminf(a, b), fminf(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7])));
      }
      const float mn = fminf(fminf(g[0], g[1]), fminf(g[2], g[3]));
      if (mn < cut && dbg_mode != 2) {
        auto ins = [&](int c) {
          if (__uint_as_float(r[c]) < cut) {
            if constexpr (REFINE) {
              if (dup_jmax != 0x7fffffff && p.perm[j0 + c] > dup_jmax) return;  // a later duplicate: cannot enter
              if (p.slices > 1) {  // the list is shared with the other slices of this block
                const int pos = atomicAdd(p.slice_cnt + t_row * 2 + half, 1);
                if (pos < TC_CAP) cl[pos] = make_uint2(r[c], static_cast<uint32_t>(j0 + c));
                return;
              }
            }
            asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(cl + cnt), "r"(r[c]), "r"(static_cast<uint32_t>(j0 + c))
                         : "memory");  // one 8-byte store
            ++cnt;
          }
        };
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (g[q] < cut) {
            if (fminf(fminf(__uint_as_float(r[8 * q + 0]), __uint_as_float(r[8 * q + 1])), __uint_as_float(r[8 * q + 2])) < cut) {
              ins(8 * q + 0);
              ins(8 * q + 1);
              ins(8 * q + 2);
            }
            if (fminf(fminf(__uint_as_float(r[8 * q + 3]), __uint_as_float(r[8 * q + 4])), __uint_as_float(r[8 * q + 5])) < cut) {
              ins(8 * q + 3);
              ins(8 * q + 4);
              ins(8 * q + 5);
            }
            if (fminf(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7])) < cut) {
              ins(8 * q + 6);
              ins(8 * q + 7);
            }
          }
        }
      }
    };

    if constexpr (WIDE) {
      // one 256-column accumulator per M block and tile: this thread's 64 columns of each 128-row half in turn
      for (int t = 0;; ++t) {
        mbar_wait(bar_acc_full + 8 * mb, static_cast<uint32_t>(t) & 1u);
        const int e = tile_seq[t % TC_RING];
        if (e == TC_SEQ_END) break;
        tc_fence_after();
        const int jt = e & ~1;
        const int nsub = (e & 1) ? 1 : 2;
        uint32_t ra[32];
        for (int sub = 0; sub < nsub; ++sub) {
          const uint32_t t_acc = t_lane + static_cast<uint32_t>(mb * 2 * TC_N + sub * TC_N + half * 64);
          tc_ld32_issue(t_acc, ra);
          tc_ld32_wait(ra);
          process(ra, jt + sub * TC_N + half * 64);
          tc_ld32_issue(t_acc + 32, ra);
          tc_ld32_wait(ra);
          if (sub == nsub - 1) {  // accumulator drained: hand it back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty + 8 * mb);
          }
          process(ra, jt + sub * TC_N + half * 64 + 32);
        }
      }
    }
    int slot = mb % NSLOT;
    uint32_t use_par = 0;  // parity of (it / NSLOT), it = 2 t + mb
    bool in_pre = pre_pass;
    int n_pre = 0;  // tiles recorded by the pre-pass: half-tile minima at [2 n + half] of the row's first list buffer
    float *pre_buf = reinterpret_cast<float *>(p.cand + t_row * 2 * TC_CAP);
    for (int t = 0; !WIDE; ++t) {
      mbar_wait(bar_acc_full + 8 * slot, use_par);
      const int jt = tile_seq[t % TC_RING];  // first (permuted) database row of the tile, or a marker
      if (jt == TC_SEQ_END) break;
      if (t == 0) stamp(2);
      tc_fence_after();
      const uint32_t t_acc = t_lane + static_cast<uint32_t>(slot * TC_N + half * 64);
      uint32_t ra[32];
      if (jt == TC_SEQ_PRE_END) {
        stamp(3);
        // end of the pre-pass: one starting threshold per row from the 2 n_pre half-tile minima both warps
        // recorded; each warp of the pair selects for 16 of the 32 rows
        if (half == 0) row_tau[prow] = tau;
        pair_sync();
        if (2 * n_pre >= p.keep) {
          // rows of this warp pair that exist: the query list is padded at its end only
          const int64_t first = q_base + mb * 128 + quad * 32;
          const int nv = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(32), p.nq - first)));
          // (at least m of a row's recorded minima are strictly below the threshold written for it)
          tc_select_threshold_rows(reinterpret_cast<const float *>(p.cand + first * 2 * TC_CAP), half * 16,
                                   min(half * 16 + 16, nv), 2 * n_pre, p.keep, lane,
                                   row_tau + mb * 128 + quad * 32);
        }
        pair_sync();
        tau = row_tau[prow];
        if (!KLOOP && p.pre_hi && active && tau < inf) {
          // the recorded minima are one-product keys: m tiles hold a one-product key below tau, hence a full key below
          // tau + (eps_single + eps_rel) (|x^| + Ymax)^2 in the worst case.  That worst case (2^-11: every dropped
          // product of every column aligned) would loosen the threshold by several key units and cost the main sweep
          // more insertions than the pre-pass saves (measured: +20 % sweep time); the dropped terms are sums of d
          // products of random sign, 2^-15 (|x^| + Ymax)^2 at one sigma, so p.pre_slack = 2^-13 is used.  A row whose
          // threshold turns out too tight ends with a short list, fails the certificate and is settled by the refine
          // pass: the starting threshold is a matter of speed, never of exactness.
          const float rr = sqrtf(xn) + p.scalars->ymax;
          tau += p.pre_slack * rr * rr;
        }
        {
          const unsigned m = __reduce_max_sync(FULL, f2ord(tau + xn));
          if (lane == 0) tau_pub[ew] = ord2f(m);
        }
        in_pre = false;
        stamp(4);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty + 8 * slot);
      } else if (in_pre) {
        // pre-pass tile: minimum key of this thread's 64 columns, nothing else
        float tmin = inf;
        auto fold = [&](const uint32_t(&r)[32]) {  // four independent chains: the dependent-issue latency is what counts
          float m4[4] = {tmin, inf, inf, inf};
#pragma unroll
          for (int q = 0; q < 32; q += 2)
            m4[(q >> 1) & 3] = fminf(fminf(__uint_as_float(r[q]), __uint_as_float(r[q + 1])), m4[(q >> 1) & 3]);
          tmin = fminf(fminf(m4[0], m4[1]), fminf(m4[2], m4[3]));
        };
        // (profiling only, results are garbage: VVB_TC_DEBUG=6 hands the accumulator straight back, 5 drains half of it)
        if (!(DBG && dbg_mode >= 6)) {
          tc_ld32_issue(t_acc, ra);
          tc_ld32_wait(ra);
          fold(ra);
        }
        if (!(DBG && dbg_mode >= 5)) {
          tc_ld32_issue(t_acc + 32, ra);
          tc_ld32_wait(ra);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty + 8 * slot);
        if (!(DBG && dbg_mode >= 5)) fold(ra);
        if (active) pre_buf[2 * n_pre + half] = tmin;
        ++n_pre;
      } else if (dbg_mode == 3) {  // profiling only: hand the accumulator straight back (MMA/TMA rate alone)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty + 8 * slot);
      } else {
        tc_ld32_issue(t_acc, ra);
        tc_ld32_wait(ra);
        process(ra, jt + half * 64);
        tc_ld32_issue(t_acc + 32, ra);
        tc_ld32_wait(ra);
        if (DBG && dbg_mode == 4)  // stress mode (VVB_TC_DEBUG=4, tests only): shake the accumulator hand-over
          __nanosleep(((static_cast<unsigned>(t) * 2654435761u + static_cast<unsigned>(ew) * 40503u) >> 20) & 0xfffu);
        // accumulator half fully drained into registers: hand it back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty + 8 * slot);
        process(ra, jt + half * 64 + 32);
      }
      // it += 2
      slot += 2;
      if (slot >= NSLOT) {
        slot -= NSLOT;
        use_par ^= 1;
      }
    }
    // final: the second half's list is appended to the first, the best m are kept, count and threshold published
    // (a sliced refine launch leaves that to refine_concat_kernel: the lists are complete only when every slice is)
    stamp(5);
    const bool deferred = !REFINE && !WIDE && p.defer_merge != 0;
    if (deferred) {  // halves_merge_kernel finishes the row
      p.slice_cnt[t_row * 2 + half] = cnt;
      p.half_tau[t_row * 2 + half] = tau;
    } else {
      half_cnt[half * TC_M + prow] = cnt;
      half_tau[half * TC_M + prow] = tau;
      pair_sync();
    }
    const bool sliced = (REFINE && p.slices > 1) || deferred;
    for (int src = half * 16; src < half * 16 + 16 && !sliced; ++src) {
      const int64_t tt = __shfl_sync(FULL, t_row, src);
      if (tt >= p.nq) continue;  // warp-uniform
      const int pr = mb * 128 + quad * 32 + src;
      const int ca = half_cnt[pr], cb = half_cnt[TC_M + pr];
      uint2 *la = p.cand + (tt + l_off) * 2 * TC_CAP, *lb = la + TC_CAP;
      // every rejected entry has key >= the threshold in force in its half when it was rejected >= that half's
      // final tau, and every row of a cell the sweep skipped has key >= both
      float thr = fminf(half_tau[pr], half_tau[TC_M + pr]);
      const int c = ca + cb;
      // keep the m smallest and at most 8 more: the re-rank kernel sorts by the exact distance anyway, so a
      // cheap threshold selection replaces a sort, and every extra candidate is a 200-byte gather there
      const int keep_max = min((p.keep + 31) & ~31, p.keep + 8);
      int nc = c;
      if (!REFINE && c > keep_max && c <= TC_CAP) {  // the common case: truncate straight from the two half lists
        if (c > 16 * TC_CAP_E) thr = fminf(thr, tc_select_truncate_halves(la, ca, lb, cb, p.keep, keep_max, lane, &nc));
        else thr = fminf(thr, tc_select_truncate_halves_short(la, ca, lb, cb, p.keep, keep_max, lane, &nc));
      } else {
        // the row's two buffers are contiguous: append the second list right behind the first (the copy may
        // run into the second buffer's own head, which has been read by then)
        for (int i0 = 0; i0 < cb; i0 += 32) {
          uint2 v = make_uint2(0u, 0u);
          if (i0 + lane < cb) v = lb[i0 + lane];
          __syncwarp();
          if (i0 + lane < cb) la[ca + i0 + lane] = v;
          __syncwarp();
        }
        // REFINE: everything below the row's threshold is handed over (c <= TC_REFINE_CAP)
        if (!REFINE && c > TC_CAP) thr = fminf(thr, tc_select_truncate_wide(la, c, p.keep, keep_max, lane, &nc));
      }
      if (lane == 0) {
        p.cand_cnt[tt + l_off] = nc;
        p.cand_thr[tt + l_off] = thr;
      }
    }
  }

  stamp(6);
  tc_fence_before();
  __syncthreads();
  stamp(7);
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// After a launch with TcParams::defer_merge: one warp per query-list entry joins the row's two half lists -- the m
// smallest keys of the union (and at most 8 more) stay, the threshold is the smaller of the halves' final thresholds and
// the truncation cut: exactly what the screening kernel's own final merge computes, at full occupancy.
template <int MINB>
__global__ void __launch_bounds__(256, MINB) halves_merge_kernel(uint2 *__restrict__ cand, const int *__restrict__ half_cnt,
                                                          const float *__restrict__ half_tau, int64_t nq, int keep,
                                                          int *__restrict__ cand_cnt, float *__restrict__ cand_thr) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (t >= nq) return;
  uint2 *la = cand + t * 2 * TC_CAP, *lb = la + TC_CAP;
  // the first 64 entries of both halves are fetched before the lengths have arrived (the buffers exist whatever the
  // lengths are): one DRAM round trip per row instead of two -- the kernel is bound by that latency
  uint2 kv[4];
  kv[0] = la[lane];
  kv[1] = la[32 + lane];
  kv[2] = lb[lane];
  kv[3] = lb[32 + lane];
  const int2 cc = *reinterpret_cast<const int2 *>(half_cnt + 2 * t);
  const float2 tt = *reinterpret_cast<const float2 *>(half_tau + 2 * t);
  const int ca = cc.x, cb = cc.y;
  float thr = fminf(tt.x, tt.y);
  const int c = ca + cb;
  const int keep_max = min((keep + 31) & ~31, keep + 8);
  int nc = c;
  if (c > keep_max && ca <= 64 && cb <= 64) {  // the common case: everything is in registers already
    const unsigned valid = (lane < ca ? 1u : 0u) | (32 + lane < ca ? 2u : 0u) | (lane < cb ? 4u : 0u) |
                           (32 + lane < cb ? 8u : 0u);
    thr = fminf(thr, warp_select_truncate_regs<4>(kv, valid, la, keep, keep_max, lane, &nc));
  } else if (c > keep_max && c <= TC_CAP) {
    if (c > 16 * TC_CAP_E) thr = fminf(thr, tc_select_truncate_halves(la, ca, lb, cb, keep, keep_max, lane, &nc));
    else thr = fminf(thr, tc_select_truncate_halves_short(la, ca, lb, cb, keep, keep_max, lane, &nc));
  } else {
    for (int i0 = 0; i0 < cb; i0 += 32) {  // (same in-place append as the screening kernel's final merge)
      uint2 v = make_uint2(0u, 0u);
      if (i0 + lane < cb) v = lb[i0 + lane];
      __syncwarp();
      if (i0 + lane < cb) la[ca + i0 + lane] = v;
      __syncwarp();
    }
    if (c > TC_CAP) thr = fminf(thr, tc_select_truncate_wide(la, c, keep, keep_max, lane, &nc));
  }
  if (lane == 0) {
    cand_cnt[t] = nc;
    cand_thr[t] = thr;
  }
}

// After a SLICED refine launch: one warp per query-list entry joins the row's two half lists (lengths in
// slice_cnt) into one and publishes its length; a half that overflowed its TC_CAP entries marks the row (thr = -inf).
__global__ void __launch_bounds__(256) refine_concat_kernel(uint2 *__restrict__ cand, const int *__restrict__ slice_cnt,
                                                           int64_t nq, int *__restrict__ cand_cnt,
                                                           float *__restrict__ cand_thr) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (t >= nq) return;
  const int ca = slice_cnt[2 * t], cb = slice_cnt[2 * t + 1];
  const bool over = ca > TC_CAP || cb > TC_CAP;
  uint2 *la = cand + t * 2 * TC_CAP, *lb = la + TC_CAP;
  if (!over)
    for (int i0 = 0; i0 < cb; i0 += 32) {  // (same in-place append as the screening kernel's final merge)
      uint2 v = make_uint2(0u, 0u);
      if (i0 + lane < cb) v = lb[i0 + lane];
      __syncwarp();
      if (i0 + lane < cb) la[ca + i0 + lane] = v;
      __syncwarp();
    }
  if (lane == 0) {
    cand_cnt[t] = over ? 0 : ca + cb;
    cand_thr[t] = over ? -__int_as_float(0x7f800000) : __int_as_float(0x7f800000);
  }
}

// After a WIDE launch with database slices: one warp per query-list entry joins the row's `wslices` lists (each at most
// keep_max long, each with its own threshold) into the row's regular list.  The m smallest keys of the union are the m
// smallest keys of the whole sweep (every slice kept its own m smallest), and every row of the database outside the
// union was rejected by some slice at a key >= that slice's threshold, so min(thresholds, truncation cuts) is the
// row's threshold.
__global__ void __launch_bounds__(256) wide_merge_kernel(const uint2 *__restrict__ cand_v, const int *__restrict__ cnt_v,
                                                        const float *__restrict__ thr_v, int wslices, int64_t nq, int keep,
                                                        int keep_max, uint2 *__restrict__ cand, int *__restrict__ cand_cnt,
                                                        float *__restrict__ cand_thr) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (t >= nq) return;
  const int64_t qb = t / TC_M, pr = t % TC_M;
  uint2 *la = cand + t * 2 * TC_CAP;
  int c = 0;
  float thr = __int_as_float(0x7f800000);
  for (int w = 0; w < wslices; ++w) {
    const int64_t vr = (qb * wslices + w) * TC_M + pr;
    const int cs = min(cnt_v[vr], TC_CAP);
    thr = fminf(thr, thr_v[vr]);
    const uint2 *src = cand_v + vr * 2 * TC_CAP;
    for (int i = lane; i < cs; i += 32) la[c + i] = src[i];
    c += cs;
    __syncwarp();
    if (c > keep_max) {  // c <= keep_max + TC_CAP <= 2 * TC_CAP here
      int nc;
      thr = fminf(thr, tc_select_truncate_wide(la, c, keep, keep_max, lane, &nc));
      c = nc;
    }
  }
  if (lane == 0) {
    cand_cnt[t] = c;
    cand_thr[t] = thr;
  }
}

// ------------------------------------------------------------------ re-rank + certificate
// One warp per query row.  E = ceil(m / 32).
// MODE 0: each lane walks its candidate's row directly.
// MODE 1: wide rows -- the 32 candidate rows of a group are staged through shared memory in 32-column
//         chunks (coalesced 128 B loads, conflict-free transposed reads).
// MODE 2: d <= 64 -- the 32 candidate rows of a group are staged WHOLE (one or two coalesced loads per
//         row from a shuffled row pointer) next to the query row; 4 warps per CTA.  The kernel is bound
//         by instruction issue (ncu: 71 % issue-active, 63 % of the warp instructions in the staging
//         loop of MODE 1 at d = 50), and this mode stages a candidate in ~7 instructions instead of ~34.
// In every mode each lane evaluates ITS candidate's chain strictly left to right, so the result is
// bit-identical.
constexpr int RR_ROW_LD = 68;  // floats per staged row: 16-byte aligned rows, and 68 = 4 (mod 32) makes the 16-byte
                               // reads of eight consecutive lanes (lane = candidate) hit all 32 banks once
template <int E, int MODE>
__global__ void __launch_bounds__(256) knn_rerank_kernel(
    const float *__restrict__ x, int64_t n, int d, int64_t q0, int64_t nq, int k, const int *__restrict__ qlist,
    const int *__restrict__ perm, const uint2 *__restrict__ cand, const int *__restrict__ cand_cnt,
    const float *__restrict__ cand_thr,
    const float *__restrict__ xnorm, const PrepScalars *__restrict__ scalars, int64_t *__restrict__ out_idx,
    float *__restrict__ out_dist, int *__restrict__ fb_count, int64_t *__restrict__ fb_rows, float eps_rel,
    float eps_refine_rel, float gamma, float *__restrict__ fb_thr, int refine) {
  constexpr bool WIDE = MODE == 1;
  constexpr bool ROWS = MODE == 2;
  __shared__ float stage[WIDE ? 8 : 1][WIDE ? 32 : 1][WIDE ? 33 : 1];
  __shared__ __align__(16) float rows_stage[ROWS ? 4 * 33 * RR_ROW_LD : 1];  // [warp][32 candidates + query][RR_ROW_LD]
  // d even and x 8-byte aligned: a row is fetched with ONE 8-byte load per lane (the kernel is bound by its shared-
  // memory and load instructions: ncu l1tex 71 %)
  const bool vec2 = ROWS && (d & 1) == 0 && (reinterpret_cast<uintptr_t>(x) & 7) == 0;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int64_t t = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (t >= nq) return;  // warp-uniform
  const int64_t g = qlist[t];  // original row of this query-list entry; candidates are permuted row ids
  const int cnt = cand_cnt[t];
  const float *xq = x + g * d;
  uint64_t key[E];
  bool self_here = false;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int pos = e * 32 + lane;
    key[e] = ~0ull;
    uint32_t j = 0;
    float acc = 0.0f;
    if (ROWS) {
      if (e * 32 < cnt) {  // warp-uniform: this group of 32 candidates is not empty
        if (pos < cnt) j = static_cast<uint32_t>(perm[cand[t * 2 * TC_CAP + pos].y]);
        const int in_group = min(32, cnt - e * 32);
        float *st = rows_stage + wib * 33 * RR_ROW_LD;
        const unsigned long long rowp = reinterpret_cast<unsigned long long>(x + static_cast<int64_t>(j) * d);
        if (vec2) {
          const bool c_ok = 2 * lane < d;
          if (e == 0)  // the query row, once
            *reinterpret_cast<float2 *>(st + 32 * RR_ROW_LD + 2 * lane) =
                c_ok ? __ldg(reinterpret_cast<const float2 *>(xq) + lane) : make_float2(0.0f, 0.0f);
          for (int cc0 = 0; cc0 < in_group; cc0 += 8) {  // 8 loads in flight, then the stores
            float2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float2 *rp = reinterpret_cast<const float2 *>(__shfl_sync(FULL, rowp, (cc0 + u) & 31));
              v[u] = (cc0 + u < in_group && c_ok) ? __ldg(rp + lane) : make_float2(0.0f, 0.0f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
              *reinterpret_cast<float2 *>(st + ((cc0 + u) & 31) * RR_ROW_LD + 2 * lane) = v[u];
          }
        } else {
          if (e == 0) {  // the query row, once
            st[32 * RR_ROW_LD + lane] = (lane < d) ? __ldg(xq + lane) : 0.0f;
            st[32 * RR_ROW_LD + 32 + lane] = (32 + lane < d) ? __ldg(xq + 32 + lane) : 0.0f;
          }
          const bool c0_ok = lane < d, c1_ok = 32 + lane < d;
          for (int cc0 = 0; cc0 < in_group; cc0 += 8) {  // 16 loads in flight, then the stores
            float v0[8], v1[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float *rp = reinterpret_cast<const float *>(__shfl_sync(FULL, rowp, (cc0 + u) & 31));
              const bool live = cc0 + u < in_group;
              v0[u] = (live && c0_ok) ? __ldg(rp + lane) : 0.0f;
              v1[u] = (live && c1_ok) ? __ldg(rp + 32 + lane) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              st[((cc0 + u) & 31) * RR_ROW_LD + lane] = v0[u];
              if (d > 32) st[((cc0 + u) & 31) * RR_ROW_LD + 32 + lane] = v1[u];
            }
          }
        }
        __syncwarp();
        // lanes beyond the group read stale rows and are ignored below.  The contract's chain: left to right, one
        // rounding per fmaf; four columns per 16-byte shared-memory read (the query's is a broadcast).
        {
          const float4 *q4 = reinterpret_cast<const float4 *>(st + 32 * RR_ROW_LD);
          const float4 *y4 = reinterpret_cast<const float4 *>(st + lane * RR_ROW_LD);
          int c = 0;
#pragma unroll 4
          for (; c + 4 <= d; c += 4) {
            const float4 a = q4[c >> 2], b = y4[c >> 2];
            float df = __fsub_rn(a.x, b.x);
            acc = __fmaf_rn(df, df, acc);
            df = __fsub_rn(a.y, b.y);
            acc = __fmaf_rn(df, df, acc);
            df = __fsub_rn(a.z, b.z);
            acc = __fmaf_rn(df, df, acc);
            df = __fsub_rn(a.w, b.w);
            acc = __fmaf_rn(df, df, acc);
          }
          for (; c < d; ++c) {
            const float df = __fsub_rn(st[32 * RR_ROW_LD + c], st[lane * RR_ROW_LD + c]);
            acc = __fmaf_rn(df, df, acc);
          }
        }
        __syncwarp();
      }
    } else if (!WIDE) {
      if (pos < cnt) {
        j = static_cast<uint32_t>(perm[cand[t * 2 * TC_CAP + pos].y]);
        const float *y = x + static_cast<int64_t>(j) * d;
        for (int c = 0; c < d; ++c) {  // the contract's chain: left to right, one rounding per fmaf
          const float df = __fsub_rn(__ldg(xq + c), __ldg(y + c));
          acc = __fmaf_rn(df, df, acc);
        }
      }
    } else {
      if (e * 32 < cnt) {  // warp-uniform: this group of 32 candidates is not empty
        if (pos < cnt) j = static_cast<uint32_t>(perm[cand[t * 2 * TC_CAP + pos].y]);
        const int in_group = min(32, cnt - e * 32);
        for (int c0 = 0; c0 < d; c0 += 32) {
          const float xv = (c0 + lane < d) ? __ldg(xq + c0 + lane) : 0.0f;
          for (int cc = 0; cc < in_group; ++cc) {
            const uint32_t jc = __shfl_sync(FULL, j, cc);
            stage[wib][cc][lane] = (c0 + lane < d) ? __ldg(x + static_cast<int64_t>(jc) * d + c0 + lane) : 0.0f;
          }
          __syncwarp();
          // every lane runs the chain (lanes beyond the group compute on stale tiles and are ignored):
          // the shuffles must be executed by the full warp
#pragma unroll
          for (int c = 0; c < 32; ++c) {  // zero-padded tail: fmaf(0, 0, acc) == acc exactly
            const float df = __fsub_rn(__shfl_sync(FULL, xv, c), stage[wib][lane][c]);
            acc = __fmaf_rn(df, df, acc);
          }
          __syncwarp();
        }
      }
    }
    if (pos < cnt) {
      if (static_cast<int64_t>(j) == g) {
        key[e] = 0ull;  // self sorts first
        self_here = true;
      } else {
        key[e] = (static_cast<uint64_t>(__float_as_uint(acc) + 1u) << 32) | j;  // d2 >= 0: bits are monotone
      }
    }
  }
  const bool has_self = __any_sync(FULL, self_here);
  warp_bitonic_sort<E>(key, lane);
  const int shift = has_self ? 0 : 1;  // self absent from the list (more than m duplicates): still column 0
  int64_t *oi = out_idx + (g - q0) * (k + 1);
  float *od = out_dist + (g - q0) * (k + 1);
  if (lane == 0) {
    oi[0] = g;
    od[0] = 0.0f;
  }
  const int pos_k = k - shift;  // sorted position of the k-th non-self neighbour
  float dk = 0.0f;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int pos = e * 32 + lane;
    const int col = pos + shift;
    const bool is_self = has_self && pos == 0;
    if (col <= k && !is_self) {
      if (pos < cnt) {
        oi[col] = static_cast<int64_t>(static_cast<uint32_t>(key[e]));
        od[col] = __fsqrt_rn(__uint_as_float(static_cast<uint32_t>(key[e] >> 32) - 1u));
      } else {
        oi[col] = -1;
        od[col] = __int_as_float(0x7fc00000);
      }
    }
    if (pos_k / 32 == e) {
      const uint64_t kv = __shfl_sync(FULL, key[e], pos_k & 31);
      dk = __uint_as_float(static_cast<uint32_t>(kv >> 32) - 1u);
    }
  }
  // certificate: no rejected row can beat the k-th neighbour, whatever the screening and the chain rounded
  const float thr = cand_thr[t];
  const float xn = xnorm[t];
  const float s = scalars->scale, ymax = scalars->ymax;
  const float r = sqrtf(xn) + ymax;
  const float eps = eps_rel * r * r;
  const float inf = __int_as_float(0x7f800000);
  bool ok = cnt > pos_k;
  if (refine) {
    // the list holds every row that can be among the k nearest (everything below the row's refine threshold, minus
    // entries beaten by k+1 others on the exact key): nothing to certify -- unless a sliced launch overflowed it
    ok = ok && thr > -inf;
  } else {
    if (ok && thr < inf) ok = (dk * s * s) < (thr + xn - eps) * (1.0f - gamma);
  }
  if (!ok && lane == 0) {
    const int slot = atomicAdd(fb_count, 1);
    fb_rows[slot] = g;
    if (!refine) {
      // refine threshold (key space): every row j whose chain distance is <= D_k has
      //   key_j <= s^2 D_k / (1 - gamma/2) - |x^|^2 + eps  <  T;   no upper bound on D_k known -> +inf
      // (eps of the REFINE sweep's keys: full precision, whatever tier rejected the row)
      float T = inf;
      if (cnt > pos_k) {
        T = fmaf(dk * s * s, 1.0f + gamma, eps_refine_rel * r * r * 1.0625f - xn);
        T += fabsf(T) * 1e-6f + 1e-30f;
      }
      fb_thr[g - q0] = T;
    }
  }
}

// ------------------------------------------------------------------ host orchestration
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// operand matrix (rows, op_cols) fp16 row-major; box = 64 columns x 128 rows, 128B swizzle
static int make_operand_map(CUtensorMap *map, void *ptr, int64_t rows, int64_t op_cols) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return VVB_ECUDA;
  }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(op_cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(op_cols) * sizeof(__half)};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ptr, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return VVB_ECUDA;
  }
  return VVB_OK;
}

// Workspace of one "session": the cell plan and the database operand (built once per call) plus the
// per-chunk query list, query operand, candidate lists and fallback list (sized for `chunk` query rows).
struct TcLayout {
  int64_t n_rows_b;  // rows of the database operand = capacity of the permuted order
  int64_t chunk_pad;
  int ka;           // K atoms per operand half
  int64_t op_cols;  // fp16 columns per operand row = 2 * ka * 64
  CellPlan cells;
  __half *opB, *opA;
  double *partial;
  float *mu;
  unsigned int *maxn2;
  PrepScalars *scalars;
  unsigned long long *tiles_done;
  int *dims;
  int *qlist;
  float *xnorm;
  uint2 *cand;
  int *cand_cnt;
  float *cand_thr;
  int *fb_count;
  int64_t *fb_rows;
  int *slice_cnt;   // (2 * chunk) list lengths of a sliced refine launch / of the halves of a deferred merge
  float *half_tau;  // (2 * chunk) final thresholds of the halves of a deferred merge
  int wslices;      // database slices of the wide-row first tier (tc_wide_slices); > 1: the three arrays below exist
  uint2 *cand_v;    // (chunk * wslices, 2 * TC_CAP) per-slice lists
  int *cand_cnt_v;  // (chunk * wslices)
  float *cand_thr_v;
  int *cta_order;   // (chunk / TC_M) launch order of the query blocks
  int *cta_scratch;
  float *fb_thr;    // (chunk) refine threshold of the rows the certificate rejected, indexed by row - q0
  void *brute_ws;
  size_t brute_ws_bytes;
  size_t total;
};

// Database slices of the wide-row first tier (TcParams::wslices): the query operands of the CTAs resident at the same
// time -- sm_count / slices blocks x 256 rows x ka x 128 bytes (hi part) -- should fill no more than about a third of
// L2, so that they AND the database tiles streaming past stay there.  VVB_TC_WSLICES forces a count (1 = off).
static int tc_wide_slices(int ka) {
  if (ka < TC_SINGLE_MIN_KA) return 1;
  int s = 1;
  if (const char *e = getenv("VVB_TC_WSLICES")) {
    s = atoi(e);
  } else {
    const double per_block = 256.0 * ka * 128.0;
    while (s < 8 && sm_count() / s * per_block > 40e6) s *= 2;
  }
  return std::max(1, std::min(s, 16));
}

static TcLayout tc_layout(void *ws, int64_t n, int d, int k, int64_t chunk) {
  TcLayout L;
  L.ka = tc_atoms(d);
  L.op_cols = 2 * static_cast<int64_t>(L.ka) * 64;
  const int P = cells_choose_count(n);
  L.n_rows_b = cells_perm_capacity(n, P);
  L.chunk_pad = (chunk + TC_M - 1) / TC_M * TC_M;
  if (L.chunk_pad == 0) L.chunk_pad = TC_M;
  Carver cv(ws);
  {
    const size_t cb = cells_carve(nullptr, nullptr, n, d, P);
    char *cbase = cv.take<char>(cb);
    cells_carve(&L.cells, cbase, n, d, P);
  }
  L.opB = cv.take<__half>(static_cast<size_t>(L.n_rows_b) * L.op_cols);
  L.opA = cv.take<__half>(static_cast<size_t>(L.chunk_pad) * L.op_cols);
  L.partial = cv.take<double>(static_cast<size_t>(PREP_BLOCKS) * L.ka * 64);
  L.mu = cv.take<float>(static_cast<size_t>(L.ka) * 64);
  L.maxn2 = cv.take<unsigned int>(1);
  L.scalars = cv.take<PrepScalars>(1);
  L.tiles_done = cv.take<unsigned long long>(2);
  L.dims = cv.take<int>(CELL_DIMS);
  L.qlist = cv.take<int>(static_cast<size_t>(L.chunk_pad));
  L.xnorm = cv.take<float>(static_cast<size_t>(L.chunk_pad));
  L.cand_cnt = cv.take<int>(static_cast<size_t>(L.chunk_pad));
  L.cand_thr = cv.take<float>(static_cast<size_t>(L.chunk_pad));
  L.fb_count = cv.take<int>(1);
  L.fb_rows = cv.take<int64_t>(static_cast<size_t>(L.chunk_pad));
  L.fb_thr = cv.take<float>(static_cast<size_t>(L.chunk_pad));
  L.slice_cnt = cv.take<int>(2 * static_cast<size_t>(L.chunk_pad));
  L.half_tau = cv.take<float>(2 * static_cast<size_t>(L.chunk_pad));
  L.cta_order = cv.take<int>(static_cast<size_t>(L.chunk_pad / TC_M));
  L.cta_scratch = cv.take<int>(cells_cta_order_scratch_ints(L.chunk_pad / TC_M));
  L.cand = cv.take<uint2>(static_cast<size_t>(L.chunk_pad) * 2 * TC_CAP);  // two lists per row (one per column half)
  L.wslices = tc_wide_slices(L.ka);
  L.cand_v = nullptr;
  L.cand_cnt_v = nullptr;
  L.cand_thr_v = nullptr;
  if (L.wslices > 1) {
    L.cand_cnt_v = cv.take<int>(static_cast<size_t>(L.chunk_pad) * L.wslices);
    L.cand_thr_v = cv.take<float>(static_cast<size_t>(L.chunk_pad) * L.wslices);
    L.cand_v = cv.take<uint2>(static_cast<size_t>(L.chunk_pad) * L.wslices * 2 * TC_CAP);
  }
  // the fallback's candidate lists reuse the screening lists (dead after the re-rank)
  L.brute_ws = L.cand;
  L.brute_ws_bytes = static_cast<size_t>(L.chunk_pad) * 2 * TC_CAP * 8;
  size_t need_brute = brute_workspace_bytes(chunk, k);
  if (knn_brute_few_workspace_bytes(k) > need_brute) need_brute = knn_brute_few_workspace_bytes(k);
  if (need_brute > L.brute_ws_bytes) {  // never for the supported k, kept for safety
    L.brute_ws = cv.take<char>(need_brute);
    L.brute_ws_bytes = need_brute;
  }
  L.total = cv.used() + 256;
  return L;
}

// Query rows processed per launch of the screening kernel.  The candidate lists cost 2 KB per row, and the
// host entry point overlaps the D2H copy of one chunk with the kernels of the next, so it uses 262,144-row
// chunks; the device entry point takes up to 1,048,576 rows at once (2 GB of lists): CTAs differ in length
// (their neighbourhoods differ in size), and one launch of 4096 CTAs has one tail where four launches of 1024
// have four (17 % of the SM time at 1M rows).
int64_t knn_tensor_chunk_rows(int64_t nq, bool host_path) {
  const int64_t cap = host_path ? 524288 : 1048576;
  return nq < cap ? nq : cap;
}

size_t knn_tensor_workspace_bytes(int64_t n, int d, int k, int64_t nq, bool host_path) {
  return tc_layout(nullptr, n, d, k, knn_tensor_chunk_rows(nq, host_path)).total;
}

// ---- session API used by api.cu ----------------------------------------------------------
struct TcSession {
  const float *x;
  int64_t n;
  int d, k, keep;
  bool single;  // first tier with one product per key (K-loop variant, ka >= TC_SINGLE_MIN_KA)
  unsigned long long tiles_seen;  // tiles_done when it was last read, and the query blocks launched since:
  double blocks_since;            // what a block of this data sweeps, for the refine-or-brute-force estimate
  TcLayout L;
  CUtensorMap map_a, map_b;
};

size_t knn_tensor_session_bytes() { return sizeof(TcSession); }

namespace {
struct ChunkEvents {  // released on every return path
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  bool on = false;
  int init() {
    for (auto &e : ev) VVB_CUDA(cudaEventCreate(&e));
    on = true;
    return VVB_OK;
  }
  ~ChunkEvents() {
    for (auto &e : ev)
      if (e) cudaEventDestroy(e);
  }
};
struct DeviceFree {
  void *p = nullptr;
  ~DeviceFree() {
    if (p) cudaFree(p);
  }
};
}  // namespace

static int tc_session_init(TcSession *S, const float *x_dev, int64_t n, int d, int k, int64_t chunk_rows,
                           void *workspace, size_t workspace_bytes) {
  VVB_REQUIRE(knn_tensor_supported(n, d, k), "tensor kNN: unsupported shape n=%lld d=%d k=%d", (long long)n, d, k);
  S->x = x_dev;
  S->n = n;
  S->d = d;
  S->k = k;
  S->keep = tc_candidates(k);
  S->tiles_seen = 0;
  S->blocks_since = 0;
  S->L = tc_layout(workspace, n, d, k, chunk_rows);
  {
    // First tier with a single product: on by default where the MMAs bound the sweep (VVB_TC_SINGLE=0/1 forces).
    // Its margin is 2^-11 of (|x^| + Ymax)^2, so it keeps more candidates per row: a larger m-th key is what lets
    // the certificate pass at that margin, and the re-rank of a wide row is cheap next to its sweep.
    const char *se = getenv("VVB_TC_SINGLE");
    S->single = S->L.ka > 1 && (se ? atoi(se) != 0 : S->L.ka >= TC_SINGLE_MIN_KA);
    if (S->single || S->L.ka >= TC_SINGLE_MIN_KA) {
      // wide rows: the margin grows with the width (and the first tier's is 2^-11), so a row needs a larger gap
      // between its k-th and m-th candidate to be certified (measured, 2M x 2000, k=50: m = 60 certifies 57 % of the
      // rows, m = 128 99.9 %)
      int m = 2 * k + 28;
      if (m > TC_CAP - 64) m = TC_CAP - 64;
      if (m > S->keep) S->keep = m;
    }
  }
  VVB_REQUIRE(workspace_bytes >= S->L.total, "tensor kNN: workspace too small (%zu < %zu)", workspace_bytes,
              S->L.total);
  VVB_CUDA(cudaFuncSetAttribute(knn_screen_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(TC_SMEM_BYTES)));
  VVB_CUDA(cudaFuncSetAttribute(knn_screen_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(TC_SMEM_BYTES_K)));
  VVB_CUDA(cudaFuncSetAttribute(knn_screen_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(TC_SMEM_BYTES)));
  VVB_CUDA(cudaFuncSetAttribute(knn_screen_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(TC_SMEM_BYTES_K)));
  VVB_CUDA(cudaFuncSetAttribute(knn_screen_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(TC_SMEM_BYTES)));
  VVB_CUDA(cudaFuncSetAttribute(knn_screen_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(TC_SMEM_BYTES_K)));
  VVB_CUDA(cudaFuncSetAttribute(knn_screen_kernel<true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(TC_SMEM_BYTES_K)));
  int rc = make_operand_map(&S->map_a, S->L.opA, S->L.chunk_pad, S->L.op_cols);
  if (rc) return rc;
  return make_operand_map(&S->map_b, S->L.opB, S->L.n_rows_b, S->L.op_cols);
}

// Preparation, phase by phase (knn_cells.cu: cells_build_phase).  Phase 0 also computes the column means, the
// largest centred norm and the geometry's coordinates (full passes over x, cheap and deterministic: every rank of a
// multi-GPU build gets the same values); phase 3 also builds the database operand (FP16 hi/lo split) in cell order.
static int tc_prepare_phase(TcSession *S, int phase, int part, int parts, cudaStream_t st) {
  TcLayout &L = S->L;
  const int64_t n = S->n;
  const int d = S->d;
  const int prep_grid = sm_count() * 8;
  int64_t row_lo = 0, row_hi = n;
  if (parts > 1) {  // the row shard of `part`: same rule as vivae_b200.sharded.shard_bounds
    const int64_t per = (n + parts - 1) / parts;
    row_lo = std::min<int64_t>(static_cast<int64_t>(part) * per, n);
    row_hi = std::min<int64_t>(row_lo + per, n);
  }
  if (phase == 0) {
    colsum_partial_kernel<<<dim3(PREP_BLOCKS, L.ka), 256, 0, st>>>(S->x, n, d, L.partial);
    VVB_CHECK_LAUNCH();
    colsum_final_kernel<<<L.ka, 64, 0, st>>>(L.partial, PREP_BLOCKS, n, d, L.mu, L.maxn2);
    VVB_CHECK_LAUNCH();
    maxnorm_kernel<<<prep_grid, 256, 0, st>>>(S->x, n, d, L.mu, L.maxn2);
    VVB_CHECK_LAUNCH();
    scalars_kernel<<<1, 1, 0, st>>>(L.maxn2, L.scalars);
    VVB_CHECK_LAUNCH();
    int rc = cells_select_dims(S->x, n, d, L.mu, L.partial, PREP_BLOCKS, L.dims, st);
    if (rc) return rc;
  }
  int rc = cells_build_phase(L.cells, S->x, d, L.mu, L.dims, reinterpret_cast<const float *>(L.scalars), phase, part,
                             parts, row_lo, row_hi, st);
  if (rc) return rc;
  if (phase == 3) {
    build_operands_kernel<false><<<prep_grid, 256, 0, st>>>(S->x, d, L.ka, L.mu, L.scalars, L.cells.perm, L.n_rows_b,
                                                            L.opB, nullptr);
    VVB_CHECK_LAUNCH();
    VVB_CUDA(cudaMemsetAsync(L.tiles_done, 0, 2 * sizeof(unsigned long long), st));
  }
  return VVB_OK;
}

// Once per call: column means, largest centred norm, the cell partition with its pairwise bounds, and
// the database operand (FP16 hi/lo split) in cell order.  `planned`: the workspace already holds all of that
// (knn_tensor_plan_phase ran phases 0..3 on it, with the ranks' exchanges in between).
int knn_tensor_begin(void *session_mem, const float *x_dev, int64_t n, int d, int k, int64_t chunk_rows,
                     void *workspace, size_t workspace_bytes, cudaStream_t st, vvb_knn_stats *stats, bool planned) {
  TcSession *S = new (session_mem) TcSession();
  int rc = tc_session_init(S, x_dev, n, d, k, chunk_rows, workspace, workspace_bytes);
  if (rc) return rc;
  ChunkEvents ce;
  if (stats) {
    rc = ce.init();
    if (rc) return rc;
    cudaEventRecord(ce.ev[0], st);
  }
  if (!planned)
    for (int phase = 0; phase < 4; ++phase) {
      rc = tc_prepare_phase(S, phase, 0, 1, st);
      if (rc) return rc;
    }
  if (stats) {
    cudaEventRecord(ce.ev[1], st);
    cudaEventSynchronize(ce.ev[1]);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ce.ev[0], ce.ev[1]);
    stats->ms_prepare += ms;
    stats->candidates = S->keep;
    stats->cells = S->L.cells.P;
    stats->products = S->single ? 1 : 3;
  }
  return VVB_OK;
}

// One phase of the preparation for part `part` of `parts` (one part per GPU), and the tables the parts must
// exchange before the next phase: xch[i] = {device pointer into the workspace, element count, dtype, op}.
//   after phase 0: cell_of (n int32)           VVB_XCH_GATHER_ROWS  every part wrote the rows of its shard
//   after phase 1: sums, sqs (P x 64 float64)  VVB_XCH_SUM          every part wrote the cells it owns, zeros elsewhere
//   after phase 2: ext (P x P), rad2 (P)       VVB_XCH_MAX_ORD      ordered-uint32 images of floats: max as UNSIGNED
int knn_tensor_plan_phase(const float *x_dev, int64_t n, int d, int k, int64_t chunk_rows, void *workspace,
                          size_t workspace_bytes, int phase, int part, int parts, cudaStream_t st, vvb_exchange *xch,
                          int *n_xch) {
  VVB_REQUIRE(phase >= 0 && phase < 4 && parts >= 1 && part >= 0 && part < parts, "knn plan: bad phase / part");
  alignas(64) char session[sizeof(TcSession)];
  TcSession *S = new (session) TcSession();
  int rc = tc_session_init(S, x_dev, n, d, k, chunk_rows, workspace, workspace_bytes);
  if (rc) return rc;
  rc = tc_prepare_phase(S, phase, part, parts, st);
  if (rc) return rc;
  const CellPlan &C = S->L.cells;
  const size_t P = static_cast<size_t>(C.P);
  int m = 0;
  if (phase == 0) {
    xch[m++] = {C.cell_of, static_cast<size_t>(n), VVB_XCH_I32, VVB_XCH_GATHER_ROWS};
  } else if (phase == 1) {
    xch[m++] = {C.sums, P * CELL_DIMS, VVB_XCH_F64, VVB_XCH_SUM};
    xch[m++] = {C.sqs, P * CELL_DIMS, VVB_XCH_F64, VVB_XCH_SUM};
  } else if (phase == 2) {
    xch[m++] = {C.ext, P * P, VVB_XCH_U32, VVB_XCH_MAX_ORD};
    xch[m++] = {C.rad2, P, VVB_XCH_U32, VVB_XCH_MAX_ORD};
  }
  *n_xch = m;
  return VVB_OK;
}

// Query rows [q0, q0+nq) (nq <= the session's chunk size): cell-sorted query list, operand, screen,
// re-rank, refine, fallback.  out_idx / out_dist point at the output row of query row q0.

int knn_tensor_chunk(void *session_mem, int64_t q0, int64_t nq, int64_t *out_idx, float *out_dist, cudaStream_t st,
                     vvb_knn_stats *stats, bool screen_only) {
  TcSession *S = static_cast<TcSession *>(session_mem);
  TcLayout &L = S->L;
  if (nq == 0) return VVB_OK;
  VVB_REQUIRE(nq <= L.chunk_pad, "tensor kNN: chunk of %lld rows exceeds the session's %lld", (long long)nq,
              (long long)L.chunk_pad);
  const int64_t nq_pad = (nq + TC_M - 1) / TC_M * TC_M;
  ChunkEvents ce;
  if (stats) {
    int erc = ce.init();
    if (erc) return erc;
  }
  auto mark = [&](int i) {
    if (stats) cudaEventRecord(ce.ev[i], st);
  };
  mark(0);
  int rc = cells_query_list(L.cells, q0, nq, nq_pad, L.qlist, nullptr, st);
  if (rc) return rc;
  build_operands_kernel<true><<<sm_count() * 8, 256, 0, st>>>(S->x, S->d, L.ka, L.mu, L.scalars, L.qlist, nq_pad, L.opA,
                                                              L.xnorm);
  VVB_CHECK_LAUNCH();
  VVB_CUDA(cudaMemsetAsync(L.fb_count, 0, sizeof(int), st));
  TcParams p;
  p.nq = nq;
  p.ks_hi = (S->d + 3 + 15) / 16;
  p.ks_lo = (S->d + 15) / 16;
  p.keep = S->keep;
  p.trigger = tc_trigger(S->keep);
  p.ka = L.ka;
  {
    const char *dbg = getenv("VVB_TC_DEBUG");
    p.debug_mode = dbg ? atoi(dbg) : 0;
  }
  p.n_cells = L.cells.P;
  p.single = S->single ? 1 : 0;
  p.eps_rel = S->single ? tc_eps_single(L.ka) : tc_eps_rel(L.ka);
  p.fixed_thr = nullptr;
  p.fixed_q0 = q0;
  p.slices = 1;
  p.wslices = 1;
  {
    const char *ph = getenv("VVB_TC_PRE_HI");
    p.pre_hi = (ph ? atoi(ph) != 0 : 1) ? 1 : 0;
    const char *ps = getenv("VVB_TC_PRE_SLACK_LOG2");  // e.g. 11 = the worst-case bound
    p.pre_slack = ldexpf(1.0f, -(ps ? atoi(ps) : 13));
  }
  p.slice_cnt = L.slice_cnt;
  p.half_tau = L.half_tau;
  {
    const char *dm = getenv("VVB_TC_DEFER_MERGE");  // 0 = the screening kernel joins the halves itself
    p.defer_merge = (dm ? atoi(dm) != 0 : 1) ? 1 : 0;
  }
  p.cta_order = nullptr;
  {
    // longest-first launch order: pays when the launch is only a few waves long (the shards of a multi-GPU build).
    // In a long launch it costs more than the tail it removes: blocks that are neighbours in the cell order sweep
    // the same database tiles at the same time, and the reordering takes that L2 locality away (1M rows, one
    // GPU, 26 waves: 21.8 -> 23.3 ms).  VVB_TC_LPT=0/1 forces it.
    const char *lpt = getenv("VVB_TC_LPT");
    const int n_blocks = static_cast<int>(nq_pad / TC_M);
    const bool want = lpt ? atoi(lpt) != 0 : (L.cells.P > 1 && n_blocks > sm_count() && n_blocks <= 8 * sm_count());
    if (want) {
      rc = cells_cta_order(L.cells, L.qlist, nq, TC_M, n_blocks, L.cta_scratch, L.cta_order, st);
      if (rc) return rc;
      p.cta_order = L.cta_order;
    }
  }
  p.x = S->x;
  p.perm = L.cells.perm;
  p.d = S->d;
  p.k1 = S->k + 1;
  p.gamma = tc_gamma(S->d);
  p.op_a = reinterpret_cast<const uint32_t *>(L.opA);
  p.qlist = L.qlist;
  p.cell_of = L.cells.cell_of;
  p.cell_off = L.cells.off;
  p.cell_lb = L.cells.lb;
  p.xnorm = L.xnorm;
  p.scalars = L.scalars;
  p.tiles_done = L.tiles_done;
  p.times = nullptr;
  const bool want_times = getenv("VVB_TC_TIMES") != nullptr;
  DeviceFree times_mem;
  if (want_times) {
    VVB_CUDA(cudaMalloc(&times_mem.p, static_cast<size_t>(nq_pad / TC_M) * 8 * sizeof(long long)));
    p.times = static_cast<long long *>(times_mem.p);
    VVB_CUDA(cudaMemsetAsync(p.times, 0, static_cast<size_t>(nq_pad / TC_M) * 8 * sizeof(long long), st));
  }
  p.cand = L.cand;
  p.cand_cnt = L.cand_cnt;
  p.cand_thr = L.cand_thr;
  const unsigned sgrid = static_cast<unsigned>(nq_pad / TC_M);
  const char *wide_env = getenv("VVB_TC_WIDE");
  const bool wide_ok = !(wide_env && atoi(wide_env) == 0);  // 256-row database tiles in the first tier (default on)
  auto launch_screen = [&](const TcParams &pp, unsigned grid) {
    const bool dbg = pp.debug_mode != 0 || pp.times != nullptr;
    if (pp.fixed_thr && L.ka == 1)
      knn_screen_kernel<false, false, true><<<dim3(grid, pp.slices), TC_THREADS, TC_SMEM_BYTES, st>>>(S->map_a, S->map_b, pp);
    else if (pp.fixed_thr)
      knn_screen_kernel<true, false, true><<<dim3(grid, pp.slices), TC_THREADS, TC_SMEM_BYTES_K, st>>>(S->map_a, S->map_b, pp);
    else if (pp.single && !dbg && wide_ok) {
      TcParams pw = pp;
      pw.wslices = L.wslices;
      if (L.wslices > 1) {
        pw.cand = L.cand_v;
        pw.cand_cnt = L.cand_cnt_v;
        pw.cand_thr = L.cand_thr_v;
      }
      knn_screen_kernel<true, false, false, true><<<grid * L.wslices, TC_THREADS, TC_SMEM_BYTES_K, st>>>(S->map_a, S->map_b, pw);
      if (L.wslices > 1) {
        const int keep_max = std::min((pp.keep + 31) & ~31, pp.keep + 8);
        wide_merge_kernel<<<static_cast<unsigned>((pp.nq + 7) / 8), 256, 0, st>>>(L.cand_v, L.cand_cnt_v, L.cand_thr_v, L.wslices,
                                                                                 pp.nq, pp.keep, keep_max, L.cand, L.cand_cnt,
                                                                                 L.cand_thr);
      }
    }
    else {
      if (L.ka == 1 && !dbg)
        knn_screen_kernel<false, false, false><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(S->map_a, S->map_b, pp);
      else if (L.ka == 1)
        knn_screen_kernel<false, true, false><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(S->map_a, S->map_b, pp);
      else if (!dbg)
        knn_screen_kernel<true, false, false><<<grid, TC_THREADS, TC_SMEM_BYTES_K, st>>>(S->map_a, S->map_b, pp);
      else
        knn_screen_kernel<true, true, false><<<grid, TC_THREADS, TC_SMEM_BYTES_K, st>>>(S->map_a, S->map_b, pp);
      if (pp.defer_merge)  // compiled for 8 CTAs per SM (32 registers): 0.97 ms at 4, 0.73 at 6, 0.63 at 8 for 1M rows
        halves_merge_kernel<8><<<static_cast<unsigned>((pp.nq + 7) / 8), 256, 0, st>>>(L.cand, L.slice_cnt, L.half_tau, pp.nq,
                                                                                      pp.keep, L.cand_cnt, L.cand_thr);
    }
  };
  mark(4);  // query list and query operand of the chunk count as preparation; ms_screen = screening launch + join of the halves
  launch_screen(p, sgrid);
  VVB_CHECK_LAUNCH();
  S->blocks_since += sgrid;
  if (want_times) {  // profiling only: mean duration of the phases of a CTA, in SM clocks
    const size_t nc = static_cast<size_t>(nq_pad / TC_M);
    std::vector<long long> h(nc * 8);
    VVB_CUDA(cudaStreamSynchronize(st));
    VVB_CUDA(cudaMemcpy(h.data(), p.times, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (size_t c = 0; c < nc; ++c)
      for (int i = 0; i < 7; ++i)
        if (h[c * 8 + i + 1] && h[c * 8 + i]) acc[i] += static_cast<double>(h[c * 8 + i + 1] - h[c * 8 + i]);
    fprintf(stderr,
            "[vvb times] CTAs %zu: prologue %.0f | first accumulator %.0f | pre-pass %.0f | pre select %.0f | main sweep "
            "%.0f | final select %.0f | exit barrier %.0f  (mean SM clocks per CTA)\n",
            nc, acc[0] / nc, acc[1] / nc, acc[2] / nc, acc[3] / nc, acc[4] / nc, acc[5] / nc, acc[6] / nc);
    p.times = nullptr;
  }
  if (screen_only) return VVB_OK;
  mark(1);
  const char *rrw = getenv("VVB_RR_WIDE_MIN");
  const int rr_wide_min = rrw ? atoi(rrw) : 65;  // d <= 64: whole-row staging (MODE 2); VVB_RR_WIDE_MIN=32 forces MODE 1
  const char *rrm = getenv("VVB_RR_ROWS");
  const bool rr_rows = S->d <= 64 && S->d < rr_wide_min && !(rrm && atoi(rrm) == 0);
  const float gamma = tc_gamma(S->d);
  // rows = query-list entries to re-rank; e_need = groups of 32 candidates a row may hold; refine = second pass
  auto launch_rerank = [&](int64_t rows, int e_need, int refine) {
    const unsigned rr_grid = static_cast<unsigned>(rr_rows ? (rows + 3) / 4 : (rows + 7) / 8);
#define VVB_RR_ARGS S->x, S->n, S->d, q0, rows, S->k, L.qlist, L.cells.perm, L.cand, L.cand_cnt, L.cand_thr, L.xnorm, \
                    L.scalars, out_idx, out_dist, L.fb_count, L.fb_rows, p.eps_rel, tc_eps_rel(L.ka), gamma, L.fb_thr, refine
#define VVB_RR(EE)                                                                            \
  do {                                                                                        \
    if (rr_rows)                                                                              \
      knn_rerank_kernel<EE, 2><<<rr_grid, 128, 0, st>>>(VVB_RR_ARGS);                         \
    else if (S->d >= 32)                                                                      \
      knn_rerank_kernel<EE, 1><<<rr_grid, 256, 0, st>>>(VVB_RR_ARGS);                         \
    else                                                                                      \
      knn_rerank_kernel<EE, 0><<<rr_grid, 256, 0, st>>>(VVB_RR_ARGS);                         \
  } while (0)
    if (e_need <= 2) VVB_RR(2);
    else if (e_need <= 4) VVB_RR(4);
    else if (e_need <= 8) VVB_RR(8);
    else VVB_RR(16);
#undef VVB_RR
#undef VVB_RR_ARGS
  };
  launch_rerank(nq, (S->keep + 31) / 32, 0);
  VVB_CHECK_LAUNCH();
  mark(2);
  // Rows the certificate rejected.  The count comes back to the host (one 4-byte copy + stream sync per chunk): it
  // sizes the refine launch, and lets a handful of brute-force rows be spread over the whole GPU (database slices)
  // instead of leaving one CTA to sweep the database alone.
  int fb = 0;
  VVB_CUDA(cudaMemcpyAsync(&fb, L.fb_count, sizeof(int), cudaMemcpyDeviceToHost, st));
  VVB_CUDA(cudaStreamSynchronize(st));
  int refined = 0;
  const char *rfe = getenv("VVB_TC_REFINE");
  bool do_refine = fb > 0 && !(rfe && atoi(rfe) == 0);
  // A refine launch of fewer blocks than SMs slices the database over the idle SMs (TcParams::slices).
  const int refine_blocks = static_cast<int>((fb + TC_M - 1) / TC_M);
  int slices = 1;
  if (fb > 0 && 2 * refine_blocks <= sm_count()) slices = std::min(sm_count() / refine_blocks, 64);
  if (const char *se = getenv("VVB_TC_REFINE_SLICES")) slices = std::max(1, std::min(atoi(se), 1024));
  if (do_refine && fb <= knn_brute_few_max_rows() && !(rfe && atoi(rfe) == 2)) {
    // A handful of rows: the refine blocks sweep their neighbourhoods on the tensor cores (MMA-bound, however few rows
    // a block holds), the few-row brute-force kernel streams the database once per row with FP32 FMAs.  Estimate
    // both (the main sweep just measured how many tiles a block of this data sweeps) and take the cheaper.
    // VVB_TC_REFINE=2 forces the refine pass.
    unsigned long long done = 0;
    VVB_CUDA(cudaMemcpy(&done, L.tiles_done, sizeof(done), cudaMemcpyDeviceToHost));
    const double tiles_per_block = static_cast<double>(done - S->tiles_seen) / std::max<double>(1.0, S->blocks_since);
    S->tiles_seen = done;
    S->blocks_since = 0;
    const double clocks_per_tile = L.ka > 1 ? 1600.0 * L.ka : 3000.0;
    const double waves = std::ceil(static_cast<double>(refine_blocks) * slices / sm_count());
    const double t_refine = waves * (tiles_per_block / slices) * clocks_per_tile / 1.9e9 + 30e-6;
    const double work = static_cast<double>(fb) * static_cast<double>(S->n) * S->d;
    const double t_brute = std::max(work / 8.0e12, work * 4.0 / 2.0e12);  // FMA rate; database bytes per row at ~2 TB/s
    if (t_brute < t_refine) do_refine = false;
  }
  if (do_refine) {
    // Refine pass (tie-aware): the failed rows become a query list of their own, swept by the same kernel with every
    // row's threshold fixed at the value the re-rank kernel derived from its k-th candidate (fb_thr); the lists,
    // the query operand and the fallback list are reused (the main pass is complete: same stream).
    refined = fb;
    const int64_t nr = fb, nr_pad = (nr + TC_M - 1) / TC_M * TC_M;
    rc = cells_query_list(L.cells, q0, nr, nr_pad, L.qlist, L.fb_rows, st);
    if (rc) return rc;
    build_operands_kernel<true><<<sm_count() * 8, 256, 0, st>>>(S->x, S->d, L.ka, L.mu, L.scalars, L.qlist, nr_pad,
                                                                L.opA, L.xnorm);
    VVB_CHECK_LAUNCH();
    VVB_CUDA(cudaMemsetAsync(L.fb_count, 0, sizeof(int), st));
    TcParams pr = p;
    pr.nq = nr;
    pr.fixed_thr = L.fb_thr;
    pr.fixed_q0 = q0;
    pr.single = 0;  // the refine sweep always runs at full precision
    pr.eps_rel = tc_eps_rel(L.ka);
    // a first tier that rejects most rows costs more than it saves: the remaining chunks skip it
    if (S->single && static_cast<int64_t>(fb) * 2 > nq) S->single = false;
    pr.debug_mode = 0;
    pr.times = nullptr;
    pr.cta_order = nullptr;
    pr.slices = slices;
    if (slices > 1) VVB_CUDA(cudaMemsetAsync(L.slice_cnt, 0, 2 * static_cast<size_t>(nr_pad) * sizeof(int), st));
    launch_screen(pr, static_cast<unsigned>(nr_pad / TC_M));
    VVB_CHECK_LAUNCH();
    if (slices > 1) {
      refine_concat_kernel<<<static_cast<unsigned>((nr + 7) / 8), 256, 0, st>>>(L.cand, L.slice_cnt, nr, L.cand_cnt,
                                                                               L.cand_thr);
      VVB_CHECK_LAUNCH();
    }
    launch_rerank(nr, TC_REFINE_CAP / 32, 1);
    VVB_CHECK_LAUNCH();
    VVB_CUDA(cudaMemcpyAsync(&fb, L.fb_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    VVB_CUDA(cudaStreamSynchronize(st));
  }
  rc = VVB_OK;
  if (fb > 0 && fb <= knn_brute_few_max_rows())
    rc = launch_knn_brute_few(S->x, S->n, S->d, S->k, L.fb_rows, fb, out_idx, out_dist, q0, L.brute_ws,
                              L.brute_ws_bytes, st);
  else if (fb > 0)
    rc = launch_knn_brute_list(S->x, S->n, S->d, S->k, L.fb_rows, L.fb_count, nq, out_idx, out_dist, q0, L.brute_ws,
                               L.brute_ws_bytes, st);
  if (rc) return rc;
  mark(3);
  if (stats) {
    VVB_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ce.ev[0], ce.ev[4]);
    stats->ms_prepare += ms;
    cudaEventElapsedTime(&ms, ce.ev[4], ce.ev[1]);
    stats->ms_screen += ms;
    cudaEventElapsedTime(&ms, ce.ev[1], ce.ev[2]);
    stats->ms_rerank += ms;
    cudaEventElapsedTime(&ms, ce.ev[2], ce.ev[3]);
    stats->ms_fallback += ms;  // refine pass + brute force
    stats->rows_fallback += fb;
    stats->rows_refined += refined;
    unsigned long long done = 0;
    int used = 0;
    VVB_CUDA(cudaMemcpy(&done, L.tiles_done, sizeof(done), cudaMemcpyDeviceToHost));
    VVB_CUDA(cudaMemcpy(&used, L.cells.off + L.cells.P, sizeof(int), cudaMemcpyDeviceToHost));
    stats->tiles_swept = static_cast<int64_t>(done);  // cumulative over the chunks of this call
    unsigned long long pre = 0;
    VVB_CUDA(cudaMemcpy(&pre, L.tiles_done + 1, sizeof(pre), cudaMemcpyDeviceToHost));
    stats->tiles_pre = static_cast<int32_t>(std::min<unsigned long long>(pre, 0x7fffffffull));
    stats->tiles_dense += static_cast<int64_t>(sgrid) * (used / TC_N);
  }
  return VVB_OK;
}

// Diagnostics: run preparation + screening only and hand the raw candidate lists back to the
// host (tests use this to measure the screening error against FP64 and to check the lists).  The
// lists are returned per ORIGINAL query row with ORIGINAL database row ids.
int knn_screen_debug(const float *x_host, int64_t n, int d, int k, int64_t q0, int64_t nq, uint32_t *cand_idx_out,
                     float *cand_key_out, int *cand_cnt_out, float *cand_thr_out, float *xnorm_out,
                     float *scale_ymax_out, float *mu_out) {
  VVB_REQUIRE(knn_tensor_supported(n, d, k) && nq >= 1 && q0 >= 0 && q0 + nq <= n, "screen debug: unsupported shape");
  VVB_REQUIRE(nq <= knn_tensor_chunk_rows(nq, true), "screen debug: at most %lld query rows",
              (long long)knn_tensor_chunk_rows(nq, true));
  const size_t ws_bytes = knn_tensor_workspace_bytes(n, d, k, nq, true);
  void *ws = nullptr, *xd = nullptr, *oi = nullptr, *od = nullptr;
  VVB_CUDA(cudaMalloc(&ws, ws_bytes));
  VVB_CUDA(cudaMalloc(&xd, static_cast<size_t>(n) * d * sizeof(float)));
  VVB_CUDA(cudaMalloc(&oi, static_cast<size_t>(nq) * (k + 1) * sizeof(int64_t)));
  VVB_CUDA(cudaMalloc(&od, static_cast<size_t>(nq) * (k + 1) * sizeof(float)));
  VVB_CUDA(cudaMemcpy(xd, x_host, static_cast<size_t>(n) * d * sizeof(float), cudaMemcpyHostToDevice));
  alignas(64) char session[sizeof(TcSession)];
  cudaStream_t st = nullptr;
  int rc = knn_tensor_begin(session, static_cast<const float *>(xd), n, d, k, nq, ws, ws_bytes, st, nullptr, false);
  if (!rc)
    rc = knn_tensor_chunk(session, q0, nq, static_cast<int64_t *>(oi), static_cast<float *>(od), st, nullptr,
                          /*screen_only=*/true);
  if (!rc) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      set_error("screen debug: %s", cudaGetErrorString(e));
      rc = VVB_ECUDA;
    }
  }
  if (!rc) {
    TcLayout &L = reinterpret_cast<TcSession *>(session)->L;
    std::vector<uint2> cl(static_cast<size_t>(nq) * 2 * TC_CAP);
    std::vector<float> ct(nq), xn(nq);
    std::vector<int> cc(nq), ql(nq), perm(static_cast<size_t>(L.n_rows_b));
    cudaMemcpy(cl.data(), L.cand, cl.size() * sizeof(uint2), cudaMemcpyDeviceToHost);
    cudaMemcpy(cc.data(), L.cand_cnt, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(ct.data(), L.cand_thr, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(xn.data(), L.xnorm, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(ql.data(), L.qlist, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(perm.data(), L.cells.perm, perm.size() * 4, cudaMemcpyDeviceToHost);
    for (int64_t t = 0; t < nq; ++t) {
      const int64_t o = ql[t] - q0;
      cand_cnt_out[o] = cc[t];
      cand_thr_out[o] = ct[t];
      xnorm_out[o] = xn[t];
      for (int e = 0; e < TC_CAP; ++e) {
        float key;
        memcpy(&key, &cl[t * 2 * TC_CAP + e].x, sizeof(float));
        cand_key_out[o * TC_CAP + e] = key;
        cand_idx_out[o * TC_CAP + e] = e < cc[t] ? static_cast<uint32_t>(perm[cl[t * 2 * TC_CAP + e].y]) : 0u;
      }
    }
    cudaMemcpy(scale_ymax_out, L.scalars, 2 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(mu_out, L.mu, static_cast<size_t>(L.ka) * 64 * sizeof(float), cudaMemcpyDeviceToHost);
  }
  cudaFree(ws);
  cudaFree(xd);
  cudaFree(oi);
  cudaFree(od);
  return rc;
}

int knn_screen_debug_capacity() { return TC_CAP; }

}  // namespace vvb
