// Micro-benchmark: TMEM read-out rate (tcgen05.ld) on sm_100a as a function of the number of
// reading warps, the load width and the number of loads in flight.  Decides whether the 64 B/clk/SM
// figure used as the roofline of the screening kernel at small d is a hardware limit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bench tmem_ld_bench.cu && ./tmem_ld_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define R4(b) "%" #b
// register-list helpers -----------------------------------------------------------------
template <int N> struct Regs { uint32_t r[N]; };

__device__ __forceinline__ void ld_x16(uint32_t a, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(a) : "memory");
}
#define L32 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}"
#define O32(r) "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
               "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), \
               "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
__device__ __forceinline__ void ld_x32(uint32_t a, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " L32 ", [%32];" : O32(r) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld_16x256b_x8(uint32_t a, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 " L32 ", [%32];" : O32(r) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld_16x128b_x16(uint32_t a, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 " L32 ", [%32];" : O32(r) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld_x32_pack(uint32_t a, uint32_t (&r)[32]) {  // 64 columns of 16-bit data -> 32 regs
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 " L32 ", [%32];" : O32(r) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: 32x32b.x32   1: 32x32b.x16   2: 16x256b.x8   3: 16x128b.x16   4: 32x32b.x64.pack16
template <int MODE, int DEPTH>
__global__ void __launch_bounds__(512, 1) bench(int reps, unsigned long long *cycles, uint32_t *sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < reps; ++it) {
    // one pass over (a 256-column half of) the lane quadrant; warps sharing a quadrant take alternating column blocks
    const int nblk = (MODE == 1) ? 32 : (MODE == 4 ? 8 : 16);   // column blocks of the 512 columns
    const int cols = 512 / nblk;
    for (int b = (warp >> 2) * DEPTH; b < nblk; b += (blockDim.x >> 7) * DEPTH) {
      if (MODE == 1) {
        uint32_t r[DEPTH][16];
#pragma unroll
        for (int q = 0; q < DEPTH; ++q) ld_x16(base + (b + q) * cols, r[q]);
        ld_wait();
#pragma unroll
        for (int q = 0; q < DEPTH; ++q) acc ^= r[q][0] ^ r[q][15];
      } else {
        uint32_t r[DEPTH][32];
#pragma unroll
        for (int q = 0; q < DEPTH; ++q) {
          const uint32_t a = base + ((b + q) % nblk) * cols;
          if (MODE == 0) ld_x32(a, r[q]);
          if (MODE == 2) { ld_16x256b_x8(a, r[q]); }
          if (MODE == 3) { ld_16x128b_x16(a, r[q]); }
          if (MODE == 4) ld_x32_pack(a, r[q]);
        }
        ld_wait();
#pragma unroll
        for (int q = 0; q < DEPTH; ++q) acc ^= r[q][0] ^ r[q][31];
      }
    }
  }
  const unsigned long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int MODE, int DEPTH>
void run(const char *name, int warps, int grid) {
  unsigned long long *cyc; uint32_t *sink;
  cudaMalloc(&cyc, grid * 8); cudaMalloc(&sink, 4);
  const int reps = 2000;
  bench<MODE, DEPTH><<<grid, warps * 32, 0>>>(10, cyc, sink);
  bench<MODE, DEPTH><<<grid, warps * 32, 0>>>(reps, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s warps=%d: %s\n", name, warps, cudaGetErrorString(e)); exit(1); }
  unsigned long long h[256]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
  unsigned long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  // bytes of TMEM cells touched per pass: the quadrant-sharing warps split the 512 columns; 4 quadrants x 32 lanes
  // (16-lane shapes touch half the lanes of the quadrant per instruction: count what was delivered to registers)
  const double regs_per_instr = (MODE == 1) ? 16 : 32;
  const int nblk = (MODE == 1) ? 32 : (MODE == 4 ? 8 : 16);
  const double instrs_per_pass_per_quadrant = nblk;   // all blocks are covered once by the warps of a quadrant
  const double bytes = 4.0 /*quadrants*/ * instrs_per_pass_per_quadrant * regs_per_instr * 32 * 4 * reps;
  printf("%-18s depth=%d warps=%2d grid=%3d: %8.1f B/clk/SM delivered to registers (%.0f cycles/pass)\n", name, DEPTH,
         warps, grid, bytes / (double)mx, (double)mx / reps);
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  for (int grid : {1, 148}) {
    for (int warps : {4, 8, 16}) {
      run<0, 1>("32x32b.x32", warps, grid);
      run<0, 2>("32x32b.x32", warps, grid);
      run<1, 2>("32x32b.x16", warps, grid);
      run<1, 4>("32x32b.x16", warps, grid);
      run<2, 1>("16x256b.x8", warps, grid);
      run<2, 2>("16x256b.x8", warps, grid);
      run<3, 2>("16x128b.x16", warps, grid);
      run<4, 1>("32x32b.x64.pack16", warps, grid);
      run<4, 2>("32x32b.x64.pack16", warps, grid);
    }
  }
  return 0;
}
