// Stand-alone probe (TEST TOOL) of the tcgen05 pieces the TMEM-operand kernel design depends on:
//   1. tcgen05.mma with the A operand in TENSOR MEMORY (written there by tcgen05.st.32x32b from the
//      thread that owns the row): lane = row, one K=16 fp16 step = 8 consecutive 32-bit columns, the
//      low half of a word = the lower k;
//   2. narrow MMAs (N = 32, on a row slice of a wider K-major no-swizzle B tile) and their cost in SM
//      cycles next to N = 128: is the tensor pipe N-proportional down to N = 32?
// Every wait is bounded.   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o ts_probe ts_probe.cu
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
  const __half* A;  // [128][K]
  const __half* B;  // [NB][K]
  float* D;         // [128][NB]
  long long* cyc;   // [4]
  int* status;
  int nslice;       // MMA N (32, 64 or 128): NB / nslice MMAs per K step
  int a_in_tmem;    // 1: A operand from TMEM, 0: from shared memory
  int reps;         // timing: repeat the whole product this many times
  int mode;         // issue mode, see the kernel
  int nacc;         // independent accumulators (copies of the product) issued round-robin per K step
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

__global__ void __launch_bounds__(128) ts_probe_kernel(args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* sB = smem;                       // [kstep][kchunk 2][n 128][16 B] = 4 KB per K step
  unsigned char* sA = smem + KSTEPS * 4096;       // same layout with 128 rows
  for (int i = tid; i < NB * K; i += 128) {
    const int n = i / K, k = i - n * K;
    const int off = (k >> 4) * 4096 + ((k >> 3) & 1) * 2048 + n * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(sB + off) = a.B[i];
    *reinterpret_cast<__half*>(sA + off) = a.A[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
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
  const uint32_t a_col0 = 448;   // A operand region: columns 256 .. 256 + 8*KSTEPS

  // ---- A -> TMEM: thread = row; K step s = 8 words at columns a_col0 + 8 s
  {
    const int row = warp * 32 + lane;
    for (int s = 0; s < KSTEPS; ++s) {
      uint32_t w[8];
      for (int i = 0; i < 8; ++i) {
        const __half lo = a.A[row * K + 16 * s + 2 * i], hi = a.A[row * K + 16 * s + 2 * i + 1];
        w[i] = (uint32_t)(*reinterpret_cast<const uint16_t*>(&lo)) | ((uint32_t)(*reinterpret_cast<const uint16_t*>(&hi)) << 16);
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + a_col0 + 8 * s;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(w[0]),
                   "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // issue: warp 1, warp-uniform control flow, operands from uniform values, MMAs under elect.sync (the way the
  // kernel issues them): measures what the tensor pipe itself sustains.  nacc accumulators are cycled round-robin.
  if (warp == 1) {
    uint32_t elected;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
    const int N = a.nslice;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sB_u = smem_u32(sB), sA_u = smem_u32(sA);
    const long long t0 = clock64();
    if (elected) {
      for (int rep = 0; rep < a.reps; ++rep) {
        for (int n0 = 0; n0 < NB; n0 += N) {
#pragma unroll
          for (int s = 0; s < KSTEPS; ++s) {
            for (int j = 0; j < a.nacc; ++j) {
              const uint64_t db = make_desc(sB_u + s * 4096 + n0 * 16, 2048, 128);
              const uint32_t acc = s > 0 ? 1u : 0u;
              const uint32_t d = tb + n0 + j * NB;
              if (a.a_in_tmem) {
                const uint32_t at = tb + a_col0 + 8 * s;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                    "r"(at), "l"(db), "r"(idesc), "r"(acc) : "memory");
              } else {
                const uint64_t da = make_desc(sA_u + s * 4096, 2048, 128);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                    "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
              }
            }
          }
        }
      }
    }
    const long long t1 = clock64();
    if (elected) {
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t ok = 0;
      for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_u32(&bar)), "r"(0)
            : "memory");
      }
      const long long t2 = clock64();
      a.cyc[0] = t1 - t0; a.cyc[1] = t2 - t0;
      if (!ok) *a.status = 1;
    }
  }
  __syncthreads();
  if (*a.status == 0) {
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 20) && !ok; ++spin) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(ok)
          : "r"(smem_u32(&bar)), "r"(0)
          : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < NB; c0 += 8) {
      uint32_t v[8];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 8; ++i) a.D[row * NB + c0 + i] = __uint_as_float(v[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

int main() {
  srand(77);
  std::vector<__half> A(128 * K), B(NB * K);
  for (auto& v : A) v = __float2half((float)(rand() % 9 - 4));
  for (auto& v : B) v = __float2half((float)(rand() % 7 - 3));
  std::vector<double> ref(128 * NB, 0.0);
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < NB; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)__half2float(A[r * K + k]) * (double)__half2float(B[n * K + k]);
      ref[r * NB + n] = s;
    }
  __half *dA, *dB;
  float* dD;
  long long* dC;
  int* dS;
  CK(cudaMalloc(&dA, A.size() * 2));
  CK(cudaMalloc(&dB, B.size() * 2));
  CK(cudaMalloc(&dD, 128 * NB * 4));
  CK(cudaMalloc(&dC, 32));
  CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  const int smem = 2 * KSTEPS * 4096 + 1024;
  CK(cudaFuncSetAttribute(ts_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int a_in_tmem = 0; a_in_tmem < 2; ++a_in_tmem)
    for (int N : {128, 64, 32})
      for (int mode : {1})
      for (int nacc : {1, 2, 3})
      for (int reps : {64}) {
        CK(cudaMemset(dD, 0xff, 128 * NB * 4));
        CK(cudaMemset(dS, 0, 4));
        args a{dA, dB, dD, dC, dS, N, a_in_tmem, reps, mode, nacc};
        ts_probe_kernel<<<1, 128, smem>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("A %s N=%d reps=%d: kernel failed: %s\n", a_in_tmem ? "tmem" : "smem", N, reps, cudaGetErrorString(e)); return 3; }
        std::vector<float> D(128 * NB);
        long long cyc[4];
        int st;
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(cyc, dC, 32, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int i = 0; i < 128 * NB; ++i) { const double err = fabs((double)D[i] - ref[i]); if (!(err <= maxerr)) maxerr = err; }
        const int n_mma = reps * (NB / N) * KSTEPS * nacc;
        printf("A in %s  N=%3d mode=%d nacc=%d reps=%2d : status=%d maxerr=%g %s | %d MMAs: issue %lld cyc, done %lld cyc -> %.1f cyc/MMA\n",
               a_in_tmem ? "TMEM" : "smem", N, mode, nacc, reps, st, maxerr, (st == 0 && maxerr == 0) ? "MATCH" : "MISMATCH", n_mma, cyc[0], cyc[1],
               (double)cyc[1] / n_mma);
      }
  printf("probe done\n");
  return 0;
}
