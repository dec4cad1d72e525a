// tcgen05 / TMA / mbarrier PTX wrappers shared by the tensor-core kernels (device only, sm_100a).
#ifndef FE_TC_CUH_
#define FE_TC_CUH_
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr uint32_t kSpinLimit = 1u << 21;  // x (<= ~2 us suspended per try) : a few seconds

// ---- PTX helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  // the suspend-time hint lets the warp sleep in hardware instead of burning issue slots
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(2000u)
      : "memory");
  return ok != 0;
}
#ifdef FE_GEMM_TRACE
__device__ int* g_fe_host_flag = nullptr;
#endif
// Bounded wait: a protocol bug turns into a reported error + trap instead of a hung GPU.
__device__ __noinline__ void mbar_timeout(int* error_flag, int code) {
#ifdef FE_GEMM_TRACE
  // debug build: every warp that times out files its code (host-mapped memory, slot = warp; the launch function prints
  // them at process exit), lingers so that the others get to file theirs, then traps
  atomicCAS(error_flag, 0, code * 1000 + (int)(threadIdx.x >> 5));
  if (g_fe_host_flag && (threadIdx.x & 31) == 0) ((volatile int*)g_fe_host_flag)[threadIdx.x >> 5] = code * 100000 + (int)blockIdx.x;
  __threadfence_system();
  for (int i = 0; i < 4000; ++i) __nanosleep(1000);
  asm volatile("trap;");
#else
  atomicExch(error_flag, code);
  __threadfence_system();
  asm volatile("trap;");
#endif
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* error_flag, int code) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < kSpinLimit; ++spin) {
    if (mbar_try_wait(bar, parity)) return;
  }
  mbar_timeout(error_flag, code);
}
// Control warps (loader, MMA issuers) wait long and often; back off between polls so their spinning does not
// take issue slots from the worker warps (30 % of all executed instructions were polls before this).
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, int* error_flag, int code) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < kSpinLimit; ++spin) {
    if (mbar_try_wait(bar, parity)) return;
    __nanosleep(128);
  }
  mbar_timeout(error_flag, code);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;   // K-chunk stride (verified on hardware: tests/cuda/umma_probe.cu)
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;   // 8-row group stride
  d |= (uint64_t)1 << 46;                        // sm_100 descriptor version
  return d;                                      // no swizzle, base offset 0
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tma_box_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Ties 8 registers loaded by an earlier tcgen05.ld to a point AFTER the wait: the compiler sees the registers as
// written here, so no use of them can be scheduled above the wait (volatile asm statements keep their order).
__device__ __forceinline__ void tmem_ld_tie8(float* v) {
  asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ int reflect_idx(int s, int T) {
  if (s < 0) s = -s;
  if (s >= T) s = 2 * (T - 1) - s;
  return s;
}


#endif  // FE_TC_CUH_
