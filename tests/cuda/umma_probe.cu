// Stand-alone probe of the tcgen05 building blocks used by the DFT-GEMM variant (TEST TOOL).
//
// One CTA computes D[128][N] = A[128][K] * B[N][K]^T with tcgen05.mma (kind::f16, fp32 accumulate in
// TMEM) from K-major, non-swizzled shared-memory operands, reads D back with tcgen05.ld and compares
// with the host.  It answers, on the real machine, the questions the kernel design depends on:
//   1. which of (LBO, SBO) in the shared-memory descriptor is the K-chunk stride and which the
//      8-row-group stride for the no-swizzle K-major canonical layout;
//   2. instruction-descriptor encoding (M=128, N in {64,128}), TMEM addressing of tcgen05.ld 32x32b;
//   3. fp16 subnormal operands are not flushed;
//   4. the accuracy of the split-fp16 (hi/lo) 3-product scheme against a float64 DFT-like product.
// Every wait is bounded, so a wrong guess reports an error instead of hanging the GPU.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tests/cuda/umma_probe.cu && ./umma_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct probe_args {
  const __half* A;   // [128][K] row-major
  const __half* B;   // [N][K] row-major
  float* D;          // [128][N]
  int* status;       // 0 ok, 1 barrier timeout
  int N, ksteps;     // K = 16 * ksteps
  int kchunk_stride; // bytes between consecutive 8-element K chunks of the same row group
  int rowgrp_stride; // bytes between consecutive 8-row groups of the same K chunk
  int swap_lbo_sbo;  // 0: LBO = kchunk_stride, SBO = rowgrp_stride ; 1: the other way round
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version for sm_100
  return d;                // layout_type = 0 (no swizzle), base_offset = 0
}

__global__ void __launch_bounds__(128) umma_probe_kernel(probe_args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = 16 * a.ksteps, N = a.N;
  const int a_bytes = 128 * K * 2;
  unsigned char* sA = smem;
  unsigned char* sB = smem + ((a_bytes + 1023) & ~1023);

  // operands -> shared memory in the layout under test
  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i - r * K;
    const int off = (k >> 3) * a.kchunk_stride + (r >> 3) * a.rowgrp_stride + (r & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(sA + off) = a.A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i - r * K;
    const int off = (k >> 3) * a.kchunk_stride + (r >> 3) * a.rowgrp_stride + (r & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(sB + off) = a.B[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 1 && lane == 0) {
    // instruction descriptor: D=f32 (bit 4), A=B=f16 (0), K-major both, N>>3 at bit 17, M>>4 at bit 24
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t lbo = a.swap_lbo_sbo ? a.rowgrp_stride : a.kchunk_stride;
    const uint32_t sbo = a.swap_lbo_sbo ? a.kchunk_stride : a.rowgrp_stride;
    for (int ks = 0; ks < a.ksteps; ++ks) {
      const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * a.kchunk_stride, lbo, sbo);
      const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * a.kchunk_stride, lbo, sbo);
      const uint32_t acc = ks > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_base),
          "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }

  // bounded wait for the MMA to finish
  uint32_t ok = 0;
  for (int spin = 0; spin < (1 << 20) && !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(&bar)), "r"(0)
        : "memory");
  }
  if (!ok) {
    if (tid == 0) *a.status = 1;
  } else {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 16; ++i) a.D[row * N + c0 + i] = __uint_as_float(v[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
}

static double run_case(const std::vector<__half>& A, const std::vector<__half>& B, int N, int ksteps, int kchunk,
                       int rowgrp, int swap, const std::vector<double>* ref_override, double* ref_scale, int* status_out) {
  const int K = 16 * ksteps;
  __half *dA, *dB;
  float* dD;
  int* dS;
  CK(cudaMalloc(&dA, A.size() * 2));
  CK(cudaMalloc(&dB, B.size() * 2));
  CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, 128 * N * 4));
  CK(cudaMemset(dS, 0, 4));
  probe_args a{dA, dB, dD, dS, N, ksteps, kchunk, rowgrp, swap};
  const int a_bytes = (128 * K * 2 + 1023) & ~1023;
  const int smem = 2 * a_bytes + 1024;  // B region sized like A: the layouts under test may leave gaps
  CK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_probe_kernel<<<1, 128, smem>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  kernel failed: %s\n", cudaGetErrorString(e));
    exit(3);
  }
  std::vector<float> D(128 * N);
  int st = 0;
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  *status_out = st;
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < N; ++n) {
      double ref;
      if (ref_override) {
        ref = (*ref_override)[r * N + n];
      } else {
        ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)__half2float(A[r * K + k]) * (double)__half2float(B[n * K + k]);
      }
      const double err = fabs((double)D[r * N + n] - ref);
      if (!(err <= maxerr)) maxerr = err;  // NaN-propagating
      if (fabs(ref) > maxref) maxref = fabs(ref);
    }
  if (ref_scale) *ref_scale = maxref;
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  return maxerr;
}

int main() {
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  srand(1234);
  // ---- 1/2: layout + descriptor hypotheses on small-integer data (exact in fp16/fp32) -------------
  for (int N : {64, 128}) {
    for (int ksteps : {1, 5}) {
      const int K = 16 * ksteps;
      std::vector<__half> A(128 * K), B(N * K);
      for (auto& v : A) v = __float2half((float)(rand() % 9 - 4));
      for (auto& v : B) v = __float2half((float)(rand() % 7 - 3));
      struct { int kchunk, rowgrp; const char* name; } layouts[] = {
          {128 * 16, 128, "chunk-major (kchunk=2048, rowgrp=128)"},
          {128, 128 * 2 * ksteps, "row-group-major (kchunk=128, rowgrp=256*ksteps)"}};
      for (auto& L : layouts) {
        if (N == 64 && L.kchunk == 2048) { /* B has 64 rows: its chunk stride would be 1024; keep 2048 (gaps) */ }
        for (int swap = 0; swap < 1; ++swap) {  // the swapped reading faults (illegal address): LBO = K-chunk stride is the right one
          int st = 0;
          double scale = 0;
          double err = run_case(A, B, N, ksteps, L.kchunk, L.rowgrp, swap, nullptr, &scale, &st);
          printf("N=%3d ksteps=%d layout=%-48s %s : status=%d maxerr=%g (ref scale %g) %s\n", N, ksteps, L.name,
                 swap ? "LBO=rowgrp,SBO=kchunk" : "LBO=kchunk,SBO=rowgrp", st, err, scale, (st == 0 && err == 0) ? "MATCH" : "");
        }
      }
    }
  }
  // ---- 3: fp16 subnormal operands ------------------------------------------------------------------
  {
    const int N = 64, ksteps = 1, K = 16;
    std::vector<__half> A(128 * K), B(N * K);
    for (auto& v : A) v = __float2half(1024.0f);
    for (auto& v : B) v = __float2half(ldexpf(1.0f, -20));  // subnormal in fp16
    for (int swap = 0; swap < 1; ++swap) {  // the swapped reading faults (illegal address): LBO = K-chunk stride is the right one
      int st = 0;
      double scale = 0;
      double err = run_case(A, B, N, ksteps, 2048, 128, swap, nullptr, &scale, &st);
      printf("subnormal test (%s): status=%d maxerr=%g expected value %g -> %s\n", swap ? "swap" : "noswap", st, err, scale,
             err == 0 ? "subnormals honoured" : "MISMATCH (flushed or wrong layout)");
    }
  }
  // ---- 4: split-fp16 accuracy on a DFT-like product ------------------------------------------------
  {
    const int N = 128, K1 = 80, ksteps = 15, K = 240;
    std::vector<float> a(128 * K1);
    std::vector<double> w(N * K1), ref(128 * N, 0.0);
    for (auto& v : a) {  // gaussian-ish samples scaled into [2^13, 2^14) like the kernel does
      float g = 0;
      for (int i = 0; i < 12; ++i) g += (float)rand() / RAND_MAX;
      v = (g - 6.0f) * 0.1f;
    }
    for (int n = 0; n < N; ++n)
      for (int j = 0; j < K1; ++j) {
        const double win = 0.5 - 0.5 * cos(2 * M_PI * (160 + 2 * j) / 320.0);
        w[n * K1 + j] = win * cos(2 * M_PI * n * (2.0 * j) / 512.0);
      }
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int j = 0; j < K1; ++j) s += (double)a[r * K1 + j] * w[n * K1 + j];
        ref[r * N + n] = s;
      }
    const float sa = ldexpf(1.0f, 14), sw = ldexpf(1.0f, 14);
    std::vector<__half> A(128 * K), B(N * K);
    for (int r = 0; r < 128; ++r)
      for (int j = 0; j < K1; ++j) {
        const float v = a[r * K1 + j] * sa;
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        A[r * K + j] = hi; A[r * K + K1 + j] = lo; A[r * K + 2 * K1 + j] = hi;
      }
    for (int n = 0; n < N; ++n)
      for (int j = 0; j < K1; ++j) {
        const double v = w[n * K1 + j] * sw;
        const __half hi = __float2half_rn((float)v);
        const __half lo = __float2half_rn((float)(v - (double)__half2float(hi)));
        B[n * K + j] = hi; B[n * K + K1 + j] = hi; B[n * K + 2 * K1 + j] = lo;
      }
    std::vector<double> ref_scaled(ref);
    for (auto& v : ref_scaled) v *= (double)sa * sw;
    for (int swap = 0; swap < 1; ++swap) {  // the swapped reading faults (illegal address): LBO = K-chunk stride is the right one
      int st = 0;
      double scale = 0;
      double err = run_case(A, B, N, ksteps, 2048, 128, swap, &ref_scaled, &scale, &st);
      printf("split-fp16 3-product accuracy (%s): status=%d max|err|/max|ref| = %.3e\n", swap ? "swap" : "noswap", st, err / scale);
    }
    // the same product with hi parts only, for contrast
    std::vector<__half> A1(128 * 80), B1(N * 80);
    for (int r = 0; r < 128; ++r) for (int j = 0; j < K1; ++j) A1[r * 80 + j] = A[r * K + j];
    for (int n = 0; n < N; ++n) for (int j = 0; j < K1; ++j) B1[n * 80 + j] = B[n * K + j];
    for (int swap = 0; swap < 1; ++swap) {  // the swapped reading faults (illegal address): LBO = K-chunk stride is the right one
      int st = 0;
      double scale = 0;
      double err = run_case(A1, B1, N, 5, 2048, 128, swap, &ref_scaled, &scale, &st);
      printf("plain fp16 1-product accuracy (%s): status=%d max|err|/max|ref| = %.3e\n", swap ? "swap" : "noswap", st, err / scale);
    }
  }
  printf("probe done\n");
  return 0;
}
