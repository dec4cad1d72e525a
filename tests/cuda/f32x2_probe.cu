// Stand-alone probe (TEST TOOL): issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on one SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_probe f32x2_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<unsigned long long*>(&d)) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return d;
}
template <int MODE>
__global__ void k(float* out, long long* cyc, const float* in) {
  float2 acc[8];
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x + i, i);
  float2 a[8], b[8];
  for (int i = 0; i < 8; ++i) { a[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + 8 + i]); b[i] = make_float2(in[threadIdx.x + 16 + i], in[threadIdx.x + 24 + i]); }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 512; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {   // 16 scalar FFMAs
        acc[i].x = fmaf(acc[i].x, a[i].x, b[i].x);
        acc[i].y = fmaf(acc[i].y, a[i].y, b[i].y);
      } else {           // 8 packed FFMA2s = the same 16 FMAs
        acc[i] = ffma2(acc[i], a[i], b[i]);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) { cyc[2 * (threadIdx.x >> 5)] = t0; cyc[2 * (threadIdx.x >> 5) + 1] = t1; }
}
int main() {
  float* d; long long* c; float* in;
  cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 1024); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
  for (int warps : {4, 8, 16}) {
    for (int mode = 0; mode < 2; ++mode) {
      if (mode == 0) k<0><<<1, warps * 32>>>(d, c, in); else k<1><<<1, warps * 32>>>(d, c, in);
      cudaDeviceSynchronize();
      long long hh[64]; cudaMemcpy(hh, c, 16 * warps, cudaMemcpyDeviceToHost);
      long long lo = hh[0], hi = hh[1]; for (int w = 0; w < warps; ++w) { if (hh[2*w] < lo) lo = hh[2*w]; if (hh[2*w+1] > hi) hi = hh[2*w+1]; }
      long long h = hi - lo;
      const double fmas = 512.0 * 16 * warps * 32;
      printf("%2d warps on one SM, %s: %lld cycles for %.0f FMAs -> %.1f FMA/clk/SM\n", warps, mode ? "FFMA2" : "FFMA ", h, fmas, fmas / h);
    }
  }
  return 0;
}
