// Stand-alone probe (TEST TOOL): tcgen05.mma.cta_group::2 — a pair of CTAs (cluster of 2) computes
// D[256][128] = A[256][K] * B[128][K]^T with M = 256 MMAs: each CTA holds its 128 rows of A and HALF of B (64 of the
// 128 N rows), the leader CTA issues, the commit is multicast to both CTAs' mbarriers, each CTA reads its own 128
// accumulator rows.  Checks the result and times the MMAs (per-SM rate against the 1-CTA form: ts_probe.cu).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o pair_probe pair_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int KSTEPS = 5, K = 16 * KSTEPS, NB = 128;

struct args {
  const __half* A;  // [256][K]
  const __half* B;  // [128][K]
  float* D;         // [256][128]
  long long* cyc;
  int* status;
  int reps, nacc;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) pair_probe_kernel(args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  unsigned char* sA = smem;                     // [kstep][kchunk 2][128 rows][16 B] = 4 KB per K step
  unsigned char* sB = smem + KSTEPS * 4096;     // [kstep][kchunk 2][64 rows][16 B]  = 2 KB per K step
  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i - r * K;
    const int off = (k >> 4) * 4096 + ((k >> 3) & 1) * 2048 + r * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(sA + off) = a.A[(rank * 128 + r) * K + k];
  }
  for (int i = tid; i < 64 * K; i += 128) {
    const int n = i / K, k = i - n * K;
    const int off = (k >> 4) * 2048 + ((k >> 3) & 1) * 1024 + n * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(sB + off) = a.B[(rank * 64 + n) * K + k];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (rank == 0 && warp == 1) {
    uint32_t elected;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
    const uint32_t idesc = (1u << 4) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sB_u = smem_u32(sB), sA_u = smem_u32(sA);
    const long long t0 = clock64();
    if (elected) {
      for (int rep = 0; rep < a.reps; ++rep) {
#pragma unroll
        for (int s = 0; s < KSTEPS; ++s) {
          for (int j = 0; j < a.nacc; ++j) {
            const uint64_t da = make_desc(sA_u + s * 4096, 2048, 128);
            const uint64_t db = make_desc(sB_u + s * 2048, 1024, 128);
            const uint32_t acc = s > 0 ? 1u : 0u;
            const uint32_t d = tb + j * NB;
            const uint32_t z = 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(d),
                "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(z) : "memory");
          }
        }
      }
    }
    const long long t1 = clock64();
    if (elected) {
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
      a.cyc[0] = t1 - t0;
    }
  }
  // both CTAs: wait for the multicast commit, read own rows
  uint32_t ok = 0;
  const long long w0 = clock64();
  for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(&bar)), "r"(0)
        : "memory");
  }
  if (rank == 0 && tid == 32) a.cyc[1] = clock64() - w0;
  if (!ok) { if (tid == 0) *a.status = 1 + rank; }
  else {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < NB; c0 += 8) {
      uint32_t v[8];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 8; ++i) a.D[(rank * 128 + row) * NB + c0 + i] = __uint_as_float(v[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

int main() {
  srand(78);
  std::vector<__half> A(256 * K), B(NB * K);
  for (auto& v : A) v = __float2half((float)(rand() % 9 - 4));
  for (auto& v : B) v = __float2half((float)(rand() % 7 - 3));
  std::vector<double> ref(256 * NB, 0.0);
  for (int r = 0; r < 256; ++r)
    for (int n = 0; n < NB; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)__half2float(A[r * K + k]) * (double)__half2float(B[n * K + k]);
      ref[r * NB + n] = s;
    }
  __half *dA, *dB; float* dD; long long* dC; int* dS;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, 256 * NB * 4));
  CK(cudaMalloc(&dC, 32)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  const int smem = KSTEPS * (4096 + 2048) + 1024;
  CK(cudaFuncSetAttribute(pair_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int nacc : {1, 2, 4})
    for (int reps : {1, 64}) {
      CK(cudaMemset(dD, 0xff, 256 * NB * 4)); CK(cudaMemset(dS, 0, 4)); CK(cudaMemset(dC, 0, 32));
      args a{dA, dB, dD, dC, dS, reps, nacc};
      pair_probe_kernel<<<2, 128, smem>>>(a);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("nacc=%d reps=%d: kernel failed: %s\n", nacc, reps, cudaGetErrorString(e)); return 3; }
      std::vector<float> D(256 * NB); long long cyc[4]; int st;
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(cyc, dC, 32, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int i = 0; i < 256 * NB; ++i) { const double err = fabs((double)D[i] - ref[i]); if (!(err <= maxerr)) maxerr = err; }
      const int n_mma = reps * KSTEPS * nacc;
      printf("cta_group::2 M=256 N=128 nacc=%d reps=%2d: status=%d maxerr=%g %s | %d MMAs: issue %lld cyc + wait %lld -> %.1f cyc/MMA (two 128-row tiles each)\n",
             nacc, reps, st, maxerr, (st == 0 && maxerr == 0) ? "MATCH" : "MISMATCH", n_mma, cyc[0], cyc[1], (double)(cyc[0] + cyc[1]) / n_mma);
    }
  printf("probe done\n");
  return 0;
}
