// DFT-GEMM variant of the spectral front-end on tcgen05 tensor cores (sm_100a).
//
//   waveform --TMA boxes--> smem samples --producer warps (fold, per-frame scale, fp16 hi/lo split)-->
//   smem A tiles --tcgen05.mma (3 split products x 4 sub-GEMMs, fp32 accumulate in TMEM)-->
//   tcgen05.ld --epilogue warps (power, sliding triangular filterbank)--> filterbank energies
//
// One persistent CTA per SM walks over tiles of 128 frames of one utterance.  Warp roles:
//   warp 0  loader   TMA tensor-map boxes of samples + cp.async.bulk of the DFT operand stage
//   warp 1  MMA      one elected thread issues tcgen05.mma / tcgen05.commit
//   warp 2  scout    per-hop-block max|x| of the next tile (per-frame power-of-two scale), L2 prefetch
//   warp 3  idle
//   warps 4-7   producers, lane <-> frame (TMEM lane): folded samples -> UMMA K-major A tiles
//   warps 8-15  epilogue, lane <-> frame; two warps per TMEM lane quarter split the columns
// Pipelines are mbarrier based (2-deep stage ring); see DESIGN.md for the protocol.
// Math, layouts and the per-thread arithmetic: fe_gemm_layout.h / fe_gemm.cuh.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fe_gemm.cuh"
#include "fe_gemm.h"
#include "fe_gemm_tables.h"

namespace {

constexpr int kThreads = 512;
constexpr int kProducerWarp0 = 4;
constexpr int kEpilogueWarp0 = 8;
constexpr int kNumEpilogueWarps = 8;
constexpr int kTileM = FE_GEMM_TILE_M;
constexpr int kSampBoxBytes = kTileM * 128;       // 128 rows x 32 floats
constexpr int kSampStageBytes = 2 * kSampBoxBytes;  // forward + backward box
constexpr int kAStageBytes = 8 * 2 * kTileM * 16;   // 32 KB
constexpr uint32_t kSpinLimit = 1u << 28;

struct gemm_args {
  const float* wave;
  const void* tables;
  float* energies;          // [rows][n_filter][n_frames]
  unsigned int* group_max;  // or NULL
  int* error_flag;          // device int, set when a barrier wait times out
  int64_t T;
  int64_t row_base;
  int32_t rows, n_frames, n_filter, hop, nhalf, nstages, kpairs;
  int32_t tiles_per_row, n_tiles, top_db_group, nb_full;
};

// ---- shared memory carve-up (offsets from a 1024-byte aligned base) -------------------------------
struct smem_layout {
  int samp, a_stage, b_stage, energies, fb, mid, bmax, unscale, p128, bars, total;
};

__host__ __device__ inline smem_layout make_layout(int nhalf, int kpairs) {
  smem_layout L;
  int off = 0;
  L.samp = off;      off += 2 * kSampStageBytes;                        // 64 KB, 1024-aligned boxes
  L.a_stage = off;   off += 2 * kAStageBytes;                           // 64 KB
  L.b_stage = off;   off += 2 * fe_gemm_b_stage_bytes(nhalf);           // 64 KB at nhalf = 128
  L.energies = off;  off += FE_GEMM_MAX_FILTERS * kTileM * 4;           // 16 KB
  L.fb = off;        off += (nhalf + 1) * (int)sizeof(fe_gemm_fb_entry);
  L.mid = off;       off += 2 * kpairs * 4;
  L.bmax = off;      off += 2 * 136 * 4;
  L.unscale = off;   off += 2 * kTileM * 4;
  L.p128 = off;      off += 2 * kTileM * 4;
  off = (off + 15) & ~15;
  L.bars = off;      off += 16 * 8;
  L.total = off;
  return L;
}

enum { BAR_SAMP_FULL = 0, BAR_STAGE_EMPTY = 2, BAR_A_FULL = 4, BAR_ACC_FULL = 6, BAR_ACC_EMPTY = 7,
       BAR_SCOUT_FULL = 8, BAR_SCOUT_EMPTY = 10, BAR_COUNT = 12 };

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
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug turns into a reported error + trap instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* error_flag, int code) {
  for (uint32_t spin = 0; spin < kSpinLimit; ++spin) {
    if (mbar_try_wait(bar, parity)) return;
  }
  atomicExch(error_flag, code);
  __threadfence_system();
  asm volatile("trap;");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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

// max |x| over [lo, hi) of one utterance, by a full warp (coalesced)
__device__ __forceinline__ float warp_absmax(const float* x, int lo, int hi, int lane) {
  float m = 0.0f;
  for (int i = lo + lane; i < hi; i += 32) m = fmaxf(m, fabsf(__ldg(x + i)));
  return warp_max(m);
}

// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
fe_gemm_kernel(const __grid_constant__ CUtensorMap wave_map, const gemm_args a) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // dynamic shared memory is only guaranteed 16-byte aligned: round up to 1024 for the swizzled boxes
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_base_s;

  const smem_layout L = make_layout(a.nhalf, a.kpairs);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned char* blob = reinterpret_cast<const unsigned char*>(a.tables);
  const fe_blob_header* h = reinterpret_cast<const fe_blob_header*>(blob);

  float* s_energy = reinterpret_cast<float*>(smem + L.energies);
  fe_gemm_fb_entry* s_fb = reinterpret_cast<fe_gemm_fb_entry*>(smem + L.fb);
  float* s_mid = reinterpret_cast<float*>(smem + L.mid);
  float* s_bmax = reinterpret_cast<float*>(smem + L.bmax);
  float* s_unscale = reinterpret_cast<float*>(smem + L.unscale);
  float* s_p128 = reinterpret_cast<float*>(smem + L.p128);
  const uint32_t bars = smem_u32(smem + L.bars);
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

  // ---- one-time setup -----------------------------------------------------------------------------
  {
    const fe_gemm_fb_entry* gfb = reinterpret_cast<const fe_gemm_fb_entry*>(blob + h->off_gemm_fb);
    for (int i = tid; i <= a.nhalf; i += kThreads) s_fb[i] = gfb[i];
    const float* gmid = reinterpret_cast<const float*>(blob + h->off_gemm_mid);
    for (int i = tid; i < 2 * a.kpairs; i += kThreads) s_mid[i] = gmid[i];
    for (int i = tid; i < FE_GEMM_MAX_FILTERS * kTileM; i += kThreads) s_energy[i] = 0.0f;
  }
  if (tid == 0) {
    mbar_init(bar(BAR_SAMP_FULL + 0), 1);
    mbar_init(bar(BAR_SAMP_FULL + 1), 1);
    mbar_init(bar(BAR_STAGE_EMPTY + 0), 1);
    mbar_init(bar(BAR_STAGE_EMPTY + 1), 1);
    mbar_init(bar(BAR_A_FULL + 0), 4);
    mbar_init(bar(BAR_A_FULL + 1), 4);
    mbar_init(bar(BAR_ACC_FULL), 1);
    mbar_init(bar(BAR_ACC_EMPTY), kNumEpilogueWarps);
    mbar_init(bar(BAR_SCOUT_FULL + 0), 1);
    mbar_init(bar(BAR_SCOUT_FULL + 1), 1);
    mbar_init(bar(BAR_SCOUT_EMPTY + 0), 4);
    mbar_init(bar(BAR_SCOUT_EMPTY + 1), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int T = (int)a.T;
  const int hop = a.hop;

  if (warp == 0) {
    // ================================ loader ==========================================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&wave_map) : "memory");
      const uint32_t b_stage_bytes = (uint32_t)fe_gemm_b_stage_bytes(a.nhalf);
      const unsigned char* gB = blob + h->off_gemm_b;
      uint32_t n = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int row_local = tile / a.tiles_per_row;
        const int t0 = (tile - row_local * a.tiles_per_row) * kTileM;
        for (int q = 0; q < a.nstages; ++q, ++n) {
          const uint32_t s = n & 1u, par = (n >> 1) & 1u;
          mbar_wait(bar(BAR_STAGE_EMPTY + s), par ^ 1u, a.error_flag, 1);
          mbar_arrive_expect_tx(bar(BAR_SAMP_FULL + s), 2u * kSampBoxBytes + b_stage_bytes);
          const uint32_t dst = smem_u32(smem + L.samp + s * kSampStageBytes);
          // forward box: block t0 + m, columns 32q .. 32q+31 ; backward box: block t0 - 1 + m, columns
          // hop - 32q - 32 .. hop - 32q - 1  (x[c - j], j = 32q+1 .. 32q+32, stored ascending in memory)
          tma_box_3d(dst, &wave_map, bar(BAR_SAMP_FULL + s), 32 * q, t0, row_local);
          tma_box_3d(dst + kSampBoxBytes, &wave_map, bar(BAR_SAMP_FULL + s), hop - 32 * q - 32, t0 - 1, row_local);
          bulk_g2s(smem_u32(smem + L.b_stage + s * b_stage_bytes), gB + (size_t)q * b_stage_bytes, b_stage_bytes,
                   bar(BAR_SAMP_FULL + s));
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ======================================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(a.nhalf >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t b_stage_bytes = (uint32_t)fe_gemm_b_stage_bytes(a.nhalf);
      const uint32_t b_lbo = (uint32_t)a.nhalf * 16u;
      uint32_t n = 0, it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        // accumulators of the previous tile must have been drained by the epilogue
        mbar_wait(bar(BAR_ACC_EMPTY), (it & 1u) ^ 1u, a.error_flag, 2);
        tc_fence_after();
        for (int q = 0; q < a.nstages; ++q, ++n) {
          const uint32_t s = n & 1u, par = (n >> 1) & 1u;
          mbar_wait(bar(BAR_SAMP_FULL + s), par, a.error_flag, 3);   // DFT operand stage landed
          mbar_wait(bar(BAR_A_FULL + s), par, a.error_flag, 4);      // producers wrote the A stage
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + L.a_stage + s * kAStageBytes);
          const uint32_t b_base = smem_u32(smem + L.b_stage + s * b_stage_bytes);
#pragma unroll
          for (int sub = 0; sub < 4; ++sub) {
            const uint64_t a_hi = make_desc(a_base + fe_gemm_a_tile_offset(sub, 0), kTileM * 16, 128);
            const uint64_t a_lo = make_desc(a_base + fe_gemm_a_tile_offset(sub, 1), kTileM * 16, 128);
            const uint64_t b_hi = make_desc(b_base + fe_gemm_b_tile_offset(a.nhalf, sub, 0), b_lbo, 128);
            const uint64_t b_lo = make_desc(b_base + fe_gemm_b_tile_offset(a.nhalf, sub, 1), b_lbo, 128);
            const uint32_t d = tmem_base + (uint32_t)(sub * a.nhalf);
            umma_f16(d, a_hi, b_hi, idesc, q > 0 ? 1u : 0u);
            umma_f16(d, a_lo, b_hi, idesc, 1u);
            umma_f16(d, a_hi, b_lo, idesc, 1u);
          }
          umma_commit(bar(BAR_STAGE_EMPTY + s));      // frees sample / A / B slot s when these MMAs retire
          if (q == a.nstages - 1) umma_commit(bar(BAR_ACC_FULL));
        }
      }
    }
  } else if (warp == 2) {
    // ================================ scout ===========================================================
    // slot s of the tile <-> hop block t0 - 1 + s, s = 0 .. 128; frame m uses slots m (backward half) and
    // m + 1 (forward half).  Blocks outside [0, nb_full) are bounded by the samples they reflect onto.
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t pb = it & 1u;
      mbar_wait(bar(BAR_SCOUT_EMPTY + pb), ((it >> 1) & 1u) ^ 1u, a.error_flag, 5);
      const int row_local = tile / a.tiles_per_row;
      const int t0 = (tile - row_local * a.tiles_per_row) * kTileM;
      const float* x = a.wave + (a.row_base + row_local) * a.T;
      float* bm = s_bmax + pb * 136;
      for (int s = 0; s <= kTileM; ++s) {
        const int b = t0 - 1 + s;
        float m;
        if (b < 0) m = warp_absmax(x, 0, min(T, 2 * hop + 1), lane);
        else if (b >= a.nb_full) m = warp_absmax(x, max(0, (a.nb_full - 2) * hop), T, lane);
        else m = warp_absmax(x, b * hop, (b + 1) * hop, lane);
        if (lane == 0) bm[s] = m;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_SCOUT_FULL + pb));
    }
  } else if (warp >= kProducerWarp0 && warp < kProducerWarp0 + 4) {
    // ================================ producers =======================================================
    const int m = (warp - kProducerWarp0) * 32 + lane;  // tile row = TMEM lane
    const float* mid_re_w = s_mid;
    const float* mid_im_w = s_mid + a.kpairs;
    uint32_t n = 0, it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t pb = it & 1u;
      const int row_local = tile / a.tiles_per_row;
      const int t0 = (tile - row_local * a.tiles_per_row) * kTileM;
      const int t = t0 + m;
      const float* x = a.wave + (a.row_base + row_local) * a.T;
      const int c = t * hop;
      // frames whose two hop blocks are not both inside the tensor map read global memory directly
      const bool edge = (t == 0) || (t >= a.nb_full);
      const bool valid = t < a.n_frames;
      mbar_wait(bar(BAR_SCOUT_FULL + pb), (it >> 1) & 1u, a.error_flag, 6);
      float scale, unscale;
      fe_gemm_frame_scale(2.0f * fmaxf(s_bmax[pb * 136 + m], s_bmax[pb * 136 + m + 1]), scale, unscale);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_SCOUT_EMPTY + pb));
      float mid_re = 0.0f, mid_im = 0.0f;
      float carry = 0.0f;  // x[c - 32q], the backward sample the previous stage's box ended with
      for (int q = 0; q < a.nstages; ++q, ++n) {
        const uint32_t s = n & 1u, par = (n >> 1) & 1u;
        mbar_wait(bar(BAR_SAMP_FULL + s), par, a.error_flag, 7);
        const unsigned char* fbox = smem + L.samp + s * kSampStageBytes + m * 128;
        const unsigned char* bbox = fbox + kSampBoxBytes;
        unsigned char* a_stage = smem + L.a_stage + s * kAStageBytes;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float fwd[16], bwd[16];
          const int j0 = 32 * q + 16 * half;
          if (!edge) {
            // swizzle-128B: 16-byte chunk ch of row m sits at chunk ch ^ (m & 7)
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const float4 f = *reinterpret_cast<const float4*>(fbox + (((4 * half + ch) ^ (m & 7)) << 4));
              fwd[4 * ch + 0] = f.x; fwd[4 * ch + 1] = f.y; fwd[4 * ch + 2] = f.z; fwd[4 * ch + 3] = f.w;
            }
            // backward box element e (0..31) = x[c - 32q - 32 + e]; bwd[i] = x[c - j0 - i]:
            //   half 0: bwd[0] = carry, bwd[i] = box[32 - i]      (i = 1..15)
            //   half 1: bwd[i] = box[16 - i]                       (i = 0..15)
            float box[20];
            const int e0 = half == 0 ? 16 : 0;  // elements e0 .. e0+15 (+ element 16 for half 1)
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const float4 f = *reinterpret_cast<const float4*>(bbox + ((((e0 >> 2) + ch) ^ (m & 7)) << 4));
              box[4 * ch + 0] = f.x; box[4 * ch + 1] = f.y; box[4 * ch + 2] = f.z; box[4 * ch + 3] = f.w;
            }
            if (half == 0) {
              bwd[0] = (q == 0) ? fwd[0] : carry;
#pragma unroll
              for (int i = 1; i < 16; ++i) bwd[i] = box[16 - i];   // element 32 - i = e0 + (16 - i)
            } else {
              const float4 f = *reinterpret_cast<const float4*>(bbox + ((4 ^ (m & 7)) << 4));  // elements 16..19
              bwd[0] = f.x;                                                                     // element 16
#pragma unroll
              for (int i = 1; i < 16; ++i) bwd[i] = box[16 - i];   // element 16 - i
              carry = box[0];                                      // element 0 = x[c - 32q - 32]
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              fwd[i] = valid ? __ldg(x + reflect_idx(c + j0 + i, T)) : 0.0f;
              bwd[i] = valid ? __ldg(x + reflect_idx(c - j0 - i, T)) : 0.0f;
            }
          }
          fe_u4 chunk[8];
          fe_gemm_produce_half(fwd, bwd, scale, j0, mid_re_w, mid_im_w, mid_re, mid_im, chunk);
#pragma unroll
          for (int sf = 0; sf < 8; ++sf) {
            *reinterpret_cast<fe_u4*>(a_stage + sf * fe_gemm_tile_bytes(kTileM) + fe_gemm_operand_offset(kTileM, m, 8 * half)) = chunk[sf];
          }
        }
        if (q == a.nstages - 1) {
          s_unscale[pb * kTileM + m] = unscale;
          s_p128[pb * kTileM + m] = fmaf(mid_re, mid_re, mid_im * mid_im);
        }
        fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_A_FULL + s));
      }
    }
  } else if (warp >= kEpilogueWarp0) {
    // ================================ epilogue ========================================================
    const int ew = warp - kEpilogueWarp0;
    const int quarter = warp & 3;          // TMEM lanes 32*quarter .. +31 are the ones this warp may read
    const int grp = ew >> 2;               // column half
    const int m = quarter * 32 + lane;
    const int kper = a.nhalf / 2;
    const int k_begin = grp * kper, k_end = k_begin + kper;
    const int etid = (warp - kEpilogueWarp0) * 32 + lane;
    const int nfil = a.n_filter;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t pb = it & 1u;
      const int row_local = tile / a.tiles_per_row;
      const int t0 = (tile - row_local * a.tiles_per_row) * kTileM;
      const int valid_rows = min(kTileM, a.n_frames - t0);
      mbar_wait(bar(BAR_ACC_FULL), it & 1u, a.error_flag, 8);
      tc_fence_after();
      const float us = s_unscale[pb * kTileM + m];
      const float p_mid = s_p128[pb * kTileM + m];
      float* ecol = s_energy + m;
      auto emit = [&](int f, float v) {
        if (f >= 0 && f < nfil && v != 0.0f) atomicAdd(ecol + f * kTileM, v * us * us);
      };
      fe_gemm_epi_state st;
      fe_gemm_epi_init(st, s_fb[k_begin]);
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16);
      for (int k0 = k_begin; k0 < k_end; k0 += 16) {
        float ce[16], co[16], se[16], so[16];
        tmem_ld16(tbase + (uint32_t)(0 * a.nhalf + k0), ce);
        tmem_ld16(tbase + (uint32_t)(1 * a.nhalf + k0), co);
        tmem_ld16(tbase + (uint32_t)(2 * a.nhalf + k0), se);
        tmem_ld16(tbase + (uint32_t)(3 * a.nhalf + k0), so);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) fe_gemm_epi_bin(st, s_fb[k0 + i], ce[i], co[i], se[i], so[i], emit);
      }
      // accumulators are in registers now: hand TMEM back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_ACC_EMPTY));
      fe_gemm_epi_flush(st, emit);
      if (grp == 1) {
        // bin n_fft/4, evaluated by the producer in true units
        const fe_gemm_fb_entry tm = s_fb[a.nhalf];
        const float p = p_mid;
        if (tm.phi_lo >= 0 && tm.phi_lo < nfil) atomicAdd(ecol + tm.phi_lo * kTileM, p * tm.w_lo_a);
        if (tm.phi_lo + 1 >= 0 && tm.phi_lo + 1 < nfil) atomicAdd(ecol + (tm.phi_lo + 1) * kTileM, p * tm.w_lo_b);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // all emissions of the tile are in s_energy
      // coalesced store: consecutive threads -> consecutive frames of one filter
      float* dst = a.energies + ((size_t)row_local * nfil) * a.n_frames + t0;
      float vmax = 0.0f;
      for (int i = etid; i < nfil * kTileM; i += kNumEpilogueWarps * 32) {
        const int f = i >> 7, r = i & (kTileM - 1);
        const float v = s_energy[i];
        s_energy[i] = 0.0f;
        if (r < valid_rows) {
          dst[(size_t)f * a.n_frames + r] = v;
          vmax = fmaxf(vmax, v);
        }
      }
      if (a.group_max) {
        vmax = warp_max(vmax);
        if (lane == 0) atomicMax(a.group_max + (a.row_base + row_local) / a.top_db_group, __float_as_uint(vmax));
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // s_energy is zero again before the next tile emits
    }
  }

  // ---- teardown -------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_fn get_encode() {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (encode_tiled_fn)p;
  }
  return fn;
}

}  // namespace

int32_t fe_gemm_compiled(void) { return 1; }

bool fe_gemm_supported(const b200fe_params* p) {
  if (p->n_filter < 1 || p->n_filter > FE_GEMM_MAX_FILTERS) return false;
  if (p->win_length != 2 * p->hop_length || p->win_length > p->n_fft) return false;
  const int kpairs = p->win_length / 2, nhalf = p->n_fft / 4;
  if (kpairs % 32 != 0 || kpairs < 32 || kpairs > 256) return false;
  if (nhalf % 16 != 0 || nhalf < 32 || nhalf > 128) return false;
  if (p->preemph != 0.0f) return false;
  return true;
}

bool fe_gemm_preferred(const b200fe_params* p) {
  (void)p;
  return false;  // until measured faster than the FFT variant (variant selection)
}
bool fe_gemm_variant_built(void) { return true; }
bool fe_gemm_auto_prefers(const b200fe_params* p) { return fe_gemm_supported(p) && fe_gemm_preferred(p); }

int64_t fe_gemm_workspace_bytes(const b200fe_params* p, int64_t chunk_rows, int64_t T) {
  (void)p; (void)chunk_rows; (void)T;
  return 256;  // error flag
}

cudaError_t fe_gemm_launch(const b200fe_params* p, const fe_fft_args& fa, int64_t row_base, int64_t rows,
                           void* gemm_ws, cudaStream_t stream, int* launches) {
  *launches = 0;
  encode_tiled_fn enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  const int hop = p->hop_length;
  const int64_t T = fa.T;
  const int nb_full = (int)(T / hop);
  gemm_args a;
  a.wave = fa.wave;
  a.tables = fa.tables;
  a.energies = fa.out;
  a.group_max = fa.group_max;
  a.error_flag = (int*)gemm_ws;
  a.T = T;
  a.row_base = row_base;
  a.rows = (int32_t)rows;
  a.n_frames = fa.n_frames;
  a.n_filter = p->n_filter;
  a.hop = hop;
  a.nhalf = p->n_fft / 4;
  a.kpairs = p->win_length / 2;
  a.nstages = a.kpairs / 32;
  a.tiles_per_row = (fa.n_frames + kTileM - 1) / kTileM;
  a.n_tiles = (int32_t)(rows * a.tiles_per_row);
  a.top_db_group = fa.top_db_group;
  a.nb_full = nb_full;

  // 3-D tensor map over the chunk's rows: [hop samples][nb_full hop blocks][rows]; box 32 x 128 x 1, swizzle 128B
  CUtensorMap map;
  const cuuint64_t gdim[3] = {(cuuint64_t)hop, (cuuint64_t)nb_full, (cuuint64_t)rows};
  const cuuint64_t gstride[2] = {(cuuint64_t)hop * 4, (cuuint64_t)T * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)kTileM, 1};
  const cuuint32_t estride[3] = {1, 1, 1};
  void* gbase = (void*)(fa.wave + row_base * T);
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, gbase, gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;

  const smem_layout L = make_layout(a.nhalf, a.kpairs);
  const int smem = L.total + 1024;
  static int attr_done = 0;
  if (attr_done < smem) {
    cudaError_t e = cudaFuncSetAttribute(fe_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done = smem;
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.n_tiles < sms ? a.n_tiles : sms;
  cudaError_t e = cudaMemsetAsync(a.error_flag, 0, 4, stream);
  if (e != cudaSuccess) return e;
  // the kernel addresses wave relative to the chunk: rows are local to the tensor map
  gemm_args b = a;
  b.wave = fa.wave;  // absolute pointer + (row_base + row_local) * T in the kernel's direct-load paths
  fe_gemm_kernel<<<grid, kThreads, smem, stream>>>(map, b);
  *launches = 1;
  return cudaGetLastError();
}
