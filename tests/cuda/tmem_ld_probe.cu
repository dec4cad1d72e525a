// Stand-alone probe (TEST TOOL): how fast can a CTA read its tensor memory back into registers?
// 16 warps (warp = TMEM lane quarter x column group of 128 columns) sweep all 512 columns x 128 lanes (256 KB) with
// tcgen05.ld.32x32b.x8 / .x16 / .x32, one tcgen05.wait::ld per `per_wait` loads.  Reports SM cycles per sweep and B/clk.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tmem_ld_probe tmem_ld_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ uint32_t ld_sum(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld_sum<8>(uint32_t taddr) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= r[i];
  return s;
}
template <>
__device__ __forceinline__ uint32_t ld_sum<16>(uint32_t taddr) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= r[i];
  return s;
}
template <>
__device__ __forceinline__ uint32_t ld_sum<32>(uint32_t taddr) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s ^= r[i];
  return s;
}

template <int X>
__global__ void __launch_bounds__(512) probe(long long* cyc, uint32_t* sink, int nwarps, int reps) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s;
  const int quarter = warp & 3, cg = warp >> 2, ngroups = nwarps / 4;
  const int cols = 512 / ngroups;
  uint32_t s = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int rep = 0; rep < reps; ++rep)
      for (int c = cg * cols; c < (cg + 1) * cols; c += X) s ^= ld_sum<X>(tb + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  sink[threadIdx.x] = s;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

int main() {
  long long* dC;
  uint32_t* dS;
  CK(cudaMalloc(&dC, 8));
  CK(cudaMalloc(&dS, 512 * 4));
  const int reps = 64;
  for (int nwarps : {4, 8, 16}) {
    for (int x : {8, 16, 32}) {
      if (x == 8) probe<8><<<1, 512>>>(dC, dS, nwarps, reps);
      else if (x == 16) probe<16><<<1, 512>>>(dC, dS, nwarps, reps);
      else probe<32><<<1, 512>>>(dC, dS, nwarps, reps);
      CK(cudaDeviceSynchronize());
      long long c;
      CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
      const double per = (double)c / reps;
      printf("%2d warps, tcgen05.ld.32x32b.x%-2d + wait each: %8.0f cycles per 256 KB sweep -> %6.1f B/clk\n", nwarps, x, per, 262144.0 / per);
    }
  }
  printf("probe done\n");
  return 0;
}
