// DFT-GEMM variant of the spectral front-end on tcgen05 tensor cores (sm_100a).
//
//   waveform --TMA boxes--> smem samples --producer warps (fold, per-frame scale, fp16 hi/lo split)-->
//   smem A tiles --tcgen05.mma (3 split products x 4 sub-GEMMs, fp32 accumulate in TMEM)-->
//   tcgen05.ld --epilogue warps (power, sliding triangular filterbank)--> filterbank energies
//
// One persistent CTA per SM walks over tiles of 128 frames of one utterance.  Warp roles:
//   warps 0-7   epilogue, lane <-> frame; two warps per TMEM lane quarter split the columns
//   warps 8-15  producers, lane <-> frame (TMEM lane): folded samples -> UMMA K-major A tiles; warps 8-11 take
//               the first 16 sample pairs of a stage, warps 12-15 the other 16
//   warp 16     loader: TMA tensor-map boxes of samples + cp.async.bulk of the DFT operand stage
//   warp 17     MMA: one elected thread issues tcgen05.mma / tcgen05.commit
//   warps 18-19 scouts: max|x| per cell of the next tile (per-frame power-of-two scale), L2 prefetch, edge frames
// Pipelines are mbarrier based (2-deep stage ring); see DESIGN.md for the protocol.
// Math, layouts and the per-thread arithmetic: fe_gemm_layout.h / fe_gemm.cuh.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fe_gemm.cuh"
#include "fe_gemm.h"
#include "fe_gemm_tables.h"

namespace {

// Warp roles.  The warp scheduler favours high warp ids, so the light, latency-critical roles (loader, MMA
// issuer, scouts) sit at the top and the heavy ones below them.
constexpr int kThreads = 640;
constexpr int kEpilogueWarp0 = 0;   // warps 0..7  (TMEM lane quarter = warp & 3)
constexpr int kNumEpilogueWarps = 8;
constexpr int kProducerWarp0 = 8;   // warps 8..15: warps 8-11 produce the first half of a stage's K range, 12-15 the second
constexpr int kNumProducerWarps = 8;
constexpr int kLoaderWarp = 16, kMmaWarp = 17;
constexpr int kScoutWarp0 = 18;     // warps 18, 19
// All roles fit the 96 registers/thread the 640-thread CTA is launched with, so no setmaxnreg rebalancing is
// done.  (Note for later: setmaxnreg.inc draws from the registers the CTA was LAUNCHED with, not from the SM's
// free registers — a budget summing to more than 96 x 640 blocks forever.)
constexpr int kTileM = FE_GEMM_TILE_M;
constexpr int kSampBoxBytes = kTileM * 128;         // 128 rows x 32 floats
constexpr int kSampStageBytes = 2 * kSampBoxBytes;  // forward + backward box
constexpr int kAStageBytes = 8 * 2 * kTileM * 16;   // 32 KB
constexpr int kCellFloats = 512;                    // scout granularity: four coalesced 512-byte warp loads
constexpr int kMaxCells = 48;
constexpr int kSideSlots = 3;                       // edge frames per tile: t = 0 and up to two at the end (slots in use: 1 + n_frames - nb_map)
constexpr uint32_t kSpinLimit = 1u << 21;  // x (<= ~2 us suspended per try) : a few seconds

struct gemm_args {
  const float* wave;
  const void* tables;
  float* energies;          // [rows][n_filter][n_frames]
  unsigned int* group_max;  // or NULL
  int* error_flag;          // device int, set when a barrier wait times out
  int64_t T;
  int64_t row_base;
  int32_t rows, n_frames, n_filter, hop, nhalf, nstages, kpairs;
  int32_t tiles_per_row, n_tiles, top_db_group, nb_map;
  int32_t fb_in_smem;  // filterbank weight table copied to shared memory (when it fits), else read through L1
};

// ---- shared memory carve-up (offsets from a 1024-byte aligned base) -------------------------------
struct smem_layout {
  int samp, a_stage, b_stage, energies, fb, ctl, mid, cells, side, side_max, unscale, p128, bars, tmem_slot, total;
};

__host__ __device__ inline smem_layout make_layout(int nhalf, int kpairs, int n_filter, int side_slots, bool fb_in_smem) {
  smem_layout L;
  int off = 0;
  L.samp = off;      off += 2 * kSampStageBytes;                        // 64 KB, 1024-aligned boxes
  L.a_stage = off;   off += 2 * kAStageBytes;                           // 64 KB
  L.b_stage = off;   off += 2 * fe_gemm_b_stage_bytes(nhalf);           // 64 KB at nhalf = 128
  L.energies = off;  off += 2 * n_filter * kTileM * 4;                  // two column groups: 20 KB at 20 filters
  L.fb = off;        off += fb_in_smem ? nhalf * (int)sizeof(fe_gemm_fbw) : 0;
  L.ctl = off;       off += (int)sizeof(fe_gemm_fbctl);
  L.mid = off;       off += 2 * kpairs * 4;
  L.cells = off;     off += 2 * kMaxCells * 4;
  L.side = off;      off += 2 * side_slots * 2 * kpairs * 4;            // 5 KB at 160 pairs, 2 slots
  L.side_max = off;  off += 2 * 4 * 4;
  // per-frame 1/scale^2 and |X[n_fft/4]|^2, producer -> epilogue.  With >= 3 stages the producers reach the
  // last stage of tile i+1 only after the epilogue of tile i has read its copy, so one buffer suffices.
  const int pe_bufs = (kpairs / FE_GEMM_STAGE_J) >= 3 ? 1 : 2;
  L.unscale = off;   off += pe_bufs * kTileM * 4;
  L.p128 = off;      off += pe_bufs * 2 * kTileM * 8;                   // per producer half: (Re, Im) of bin n_fft/4
  off = (off + 15) & ~15;
  L.bars = off;      off += 16 * 8;
  L.tmem_slot = off; off += 16;
  L.total = off;
  return L;
}

enum { BAR_SAMP_FULL = 0, BAR_STAGE_EMPTY = 2, BAR_A_FULL = 4, BAR_ACC_FULL = 6, BAR_ACC_EMPTY = 7,
       BAR_SCOUT_FULL = 8, BAR_SCOUT_EMPTY = 10, BAR_SAMP_EMPTY = 12, BAR_B_FULL = 14, BAR_COUNT = 16 };

#ifdef FE_GEMM_TRACE
// debug build only: SM-clock timestamps of pipeline events of CTA 0 into the (enlarged) error-flag buffer
#define FE_TRACE(ev, it, q) do { if (blockIdx.x == 0 && (it) < 8) { ((long long*)(a.error_flag + 64))[((it) * 8 + (q)) * 16 + (ev)] = clock64(); } } while (0)
#else
#define FE_TRACE(ev, it, q) do { } while (0)
#endif

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
// Bounded wait: a protocol bug turns into a reported error + trap instead of a hung GPU.
__device__ __noinline__ void mbar_timeout(int* error_flag, int code) {
#ifdef FE_GEMM_TRACE
  // debug build: record who timed out first (code * 1000 + warp) and leave, so the flag can be read back
  atomicCAS(error_flag, 0, code * 1000 + (int)(threadIdx.x >> 5));
  __threadfence_system();
  asm volatile("exit;");
#else
  atomicExch(error_flag, code);
  __threadfence_system();
  asm volatile("trap;");
#endif
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* error_flag, int code) {
  for (uint32_t spin = 0; spin < kSpinLimit; ++spin) {
    if (mbar_try_wait(bar, parity)) return;
  }
  mbar_timeout(error_flag, code);
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
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

// edge-frame slot of frame t within its tile: 0 for t == 0, 1.. for frames whose forward block is not
// covered by the tensor map (t >= nb_map); -1 for ordinary frames
__device__ __forceinline__ int edge_slot(int t, int nb_map, int n_frames) {
  if (t == 0) return 0;
  if (t >= nb_map && t < n_frames) return 1 + (t - nb_map);
  return -1;
}

// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
fe_gemm_kernel(const __grid_constant__ CUtensorMap wave_map, const gemm_args a) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // the swizzled TMA boxes need a 1024-byte aligned window; __align__(1024) on the extern array is honoured
  // (no static shared memory in this kernel) — checked, not assumed
  unsigned char* smem = smem_dyn;
  if ((smem_u32(smem_dyn) & 1023u) != 0u) {
    if (threadIdx.x == 0) atomicExch(a.error_flag, 99);
    return;
  }

  const smem_layout L = make_layout(a.nhalf, a.kpairs, a.n_filter, 1 + a.n_frames - a.nb_map, a.fb_in_smem != 0);
  const int side_slots = 1 + a.n_frames - a.nb_map;
  const int pe_stride = a.nstages >= 3 ? 0 : kTileM;  // see make_layout
  const int side_floats = 2 * a.kpairs;  // per edge slot: forward[kpairs] + backward in box order [kpairs]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned char* blob = reinterpret_cast<const unsigned char*>(a.tables);
  const fe_blob_header* h = reinterpret_cast<const fe_blob_header*>(blob);

  float* s_energy = reinterpret_cast<float*>(smem + L.energies);
  const fe_gemm_fbw* g_fb = reinterpret_cast<const fe_gemm_fbw*>(blob + h->off_gemm_fb);
  fe_gemm_fbw* s_fb = reinterpret_cast<fe_gemm_fbw*>(smem + L.fb);
  fe_gemm_fbctl* s_ctl = reinterpret_cast<fe_gemm_fbctl*>(smem + L.ctl);
  float* s_mid = reinterpret_cast<float*>(smem + L.mid);
  float* s_cells = reinterpret_cast<float*>(smem + L.cells);
  float* s_side = reinterpret_cast<float*>(smem + L.side);
  float* s_side_max = reinterpret_cast<float*>(smem + L.side_max);
  float* s_unscale = reinterpret_cast<float*>(smem + L.unscale);
  float2* s_p128 = reinterpret_cast<float2*>(smem + L.p128);  // [buffer][half][frame]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L.tmem_slot);
  const uint32_t bars = smem_u32(smem + L.bars);
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

  // ---- one-time setup -----------------------------------------------------------------------------
  {
    if (a.fb_in_smem) for (int i = tid; i < a.nhalf; i += kThreads) s_fb[i] = g_fb[i];
    const int32_t* gctl = reinterpret_cast<const int32_t*>(blob + h->off_gemm_fbflag);
    for (int i = tid; i < (int)(sizeof(fe_gemm_fbctl) / 4); i += kThreads) reinterpret_cast<int32_t*>(s_ctl)[i] = gctl[i];
    const float* gmid = reinterpret_cast<const float*>(blob + h->off_gemm_mid);
    for (int i = tid; i < 2 * a.kpairs; i += kThreads) s_mid[i] = gmid[i];
    for (int i = tid; i < 2 * a.n_filter * kTileM; i += kThreads) s_energy[i] = 0.0f;
  }
  if (tid == 0) {
    mbar_init(bar(BAR_SAMP_FULL + 0), 1);
    mbar_init(bar(BAR_SAMP_FULL + 1), 1);
    mbar_init(bar(BAR_STAGE_EMPTY + 0), 1);
    mbar_init(bar(BAR_STAGE_EMPTY + 1), 1);
    mbar_init(bar(BAR_A_FULL + 0), kNumProducerWarps);
    mbar_init(bar(BAR_A_FULL + 1), kNumProducerWarps);
    mbar_init(bar(BAR_ACC_FULL), 1);
    mbar_init(bar(BAR_ACC_EMPTY), kNumEpilogueWarps);
    mbar_init(bar(BAR_SCOUT_FULL + 0), 2);
    mbar_init(bar(BAR_SCOUT_FULL + 1), 2);
    mbar_init(bar(BAR_SCOUT_EMPTY + 0), kNumProducerWarps);
    mbar_init(bar(BAR_SCOUT_EMPTY + 1), kNumProducerWarps);
    mbar_init(bar(BAR_SAMP_EMPTY + 0), kNumProducerWarps);
    mbar_init(bar(BAR_SAMP_EMPTY + 1), kNumProducerWarps);
    mbar_init(bar(BAR_B_FULL + 0), 1);
    mbar_init(bar(BAR_B_FULL + 1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int T = (int)a.T;
  const int hop = a.hop;

  if (warp == kLoaderWarp) {
    // ================================ loader ==========================================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&wave_map) : "memory");
      const uint32_t b_stage_bytes = (uint32_t)fe_gemm_b_stage_bytes(a.nhalf);
      const unsigned char* gB = blob + h->off_gemm_b;
      uint32_t n = 0, lit = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++lit) {
        const int row_local = tile / a.tiles_per_row;
        const int t0 = (tile - row_local * a.tiles_per_row) * kTileM;
        for (int q = 0; q < a.nstages; ++q, ++n) {
          const uint32_t s = n & 1u, par = (n >> 1) & 1u;
          // the sample boxes of slot s are free once the producers have read them (two stages ago) ...
          mbar_wait(bar(BAR_SAMP_EMPTY + s), par ^ 1u, a.error_flag, 1);
          FE_TRACE(0, lit, q);
          mbar_arrive_expect_tx(bar(BAR_SAMP_FULL + s), 2u * kSampBoxBytes);
          const uint32_t dst = smem_u32(smem + L.samp + s * kSampStageBytes);
          // forward box : block t0 + m,     columns 32q .. 32q+31        = x[c + 32q + e]
          // backward box: block t0 - 1 + m, columns hop-32q-32 .. hop-32q-1 = x[c - 32q - 32 + e]
          //               (16-byte aligned columns: an odd start column made the TMA fault)
          tma_box_3d(dst, &wave_map, bar(BAR_SAMP_FULL + s), 32 * q, t0, row_local);
          tma_box_3d(dst + kSampBoxBytes, &wave_map, bar(BAR_SAMP_FULL + s), hop - 32 * q - 32, t0 - 1, row_local);
          // ... the DFT operand slot only when the MMAs that read it have retired
          mbar_wait(bar(BAR_STAGE_EMPTY + s), par ^ 1u, a.error_flag, 9);
          mbar_arrive_expect_tx(bar(BAR_B_FULL + s), b_stage_bytes);
          bulk_g2s(smem_u32(smem + L.b_stage + s * b_stage_bytes), gB + (size_t)q * b_stage_bytes, b_stage_bytes,
                   bar(BAR_B_FULL + s));
        }
      }
    }
  } else if (warp == kMmaWarp) {
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
        FE_TRACE(7, it, 0);
        for (int q = 0; q < a.nstages; ++q, ++n) {
          const uint32_t s = n & 1u, par = (n >> 1) & 1u;
          mbar_wait(bar(BAR_B_FULL + s), par, a.error_flag, 3);      // DFT operand stage landed
          mbar_wait(bar(BAR_A_FULL + s), par, a.error_flag, 4);      // producers wrote the A stage
          tc_fence_after();
          FE_TRACE(3, it, q);
          const uint32_t a_base = smem_u32(smem + L.a_stage + s * kAStageBytes);
          const uint32_t b_base = smem_u32(smem + L.b_stage + s * b_stage_bytes);
#pragma unroll 1
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
  } else if (warp >= kScoutWarp0) {
    // ================================ scouts (warps 18, 19) =============================================
    // Per-tile max|x| over 128-float cells (one coalesced warp load each) for the per-frame fp16 scale,
    // one tile ahead of the producers; the reads also pull the tile's samples into L2 ahead of the TMA.
    // Warp 2 additionally gathers the reflect-padded samples of edge frames into side buffers.
    const int sw = warp - kScoutWarp0;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t pb = it & 1u;
      mbar_wait(bar(BAR_SCOUT_EMPTY + pb), ((it >> 1) & 1u) ^ 1u, a.error_flag, 5);
      if (sw == 0 && lane == 0) FE_TRACE(10, it, 0);
      const int row_local = tile / a.tiles_per_row;
      const int t0 = (tile - row_local * a.tiles_per_row) * kTileM;
      const float* x = a.wave + (a.row_base + row_local) * a.T;
      float* cells = s_cells + pb * kMaxCells;
      const int c_first = max(0, (t0 - 1) * hop) / kCellFloats;
      const int c_last = (min(T, (t0 + kTileM) * hop) - 1) / kCellFloats;
      {
        // pull the NEXT tile's samples into L2 now: its scout loads and TMA boxes then hit L2
        const int ntile = tile + gridDim.x;
        if (ntile < a.n_tiles && sw == 0) {
          const int nrow = ntile / a.tiles_per_row;
          const int nt0 = (ntile - nrow * a.tiles_per_row) * kTileM;
          const int lo = (max(0, (nt0 - 1) * hop) / kCellFloats) * kCellFloats;
          const int hi = min(T, (nt0 + kTileM) * hop);
          const float* nx = a.wave + (a.row_base + nrow) * a.T;
          for (int off = lo + lane * 1024; off < hi; off += 32 * 1024) {   // 4 KB pieces, one per lane
            const int nbytes = min(1024, hi - off) * 4;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nx + off), "r"(nbytes) : "memory");
          }
        }
      }
      // each scout warp takes every other group of 4 cells: 16 independent 16-byte loads per lane in flight,
      // then one lane-local max and one shuffle reduction per cell
      for (int c0 = c_first + 4 * sw; c0 <= c_last; c0 += 8) {
        float4 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int c = c0 + (u >> 2);
          const int idx = c * kCellFloats + (u & 3) * 128 + 4 * lane;
          v[u] = (c <= c_last && idx < T) ? __ldg(reinterpret_cast<const float4*>(x + idx)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float mx = 0.0f;
#pragma unroll
          for (int u = 4 * g; u < 4 * g + 4; ++u)
            mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w))));
          mx = warp_max(mx);
          if (lane == 0 && c0 + g <= c_last) cells[c0 + g - c_first] = mx;
        }
      }
      if (sw == 0) {
        for (int m = 0; m < kTileM; ++m) {
          const int t = t0 + m;
          const int slot = edge_slot(t, a.nb_map, a.n_frames);
          if (slot < 0) {
            if (t >= 1 && t < a.nb_map) { m = min(kTileM, a.nb_map - t0) - 1; }  // skip the ordinary frames
            continue;
          }
          float* sf = s_side + (pb * side_slots + slot) * side_floats;
          float* sb = sf + a.kpairs;
          const int c = t * hop;
          float mx = 0.0f;
          for (int j = lane; j < a.kpairs; j += 32) {
            const float f = __ldg(x + reflect_idx(c + j, T));
            const int q = j >> 5, e = j & 31;
            const float b = __ldg(x + reflect_idx(c - 32 * q - 32 + e, T));
            sf[j] = f;
            sb[j] = b;
            mx = fmaxf(mx, fmaxf(fabsf(f), fabsf(b)));
          }
          mx = warp_max(mx);
          if (lane == 0) s_side_max[pb * 4 + slot] = mx;
        }
      }
      __syncwarp();
      if (sw == 0 && lane == 0) FE_TRACE(11, it, 0);
      if (lane == 0) mbar_arrive(bar(BAR_SCOUT_FULL + pb));
    }
  } else if (warp >= kProducerWarp0) {
    // ================================ producers (warps 8..15) =========================================
    // warp pair (w, w+4) shares 32 frames: warps 4-7 produce sample pairs j = 32q .. 32q+15 of every stage
    // (K chunk 0 of the A tiles), warps 8-11 j = 32q+16 .. 32q+31 (K chunk 1)
    const int pw = warp - kProducerWarp0;
    const int half = pw >> 2;
    const int m = (pw & 3) * 32 + lane;  // tile row = TMEM lane
    const float* mid_re_w = s_mid;
    const float* mid_im_w = s_mid + a.kpairs;
    const int swz = m & 7;
    uint32_t n = 0, it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t pb = it & 1u;
      const int row_local = tile / a.tiles_per_row;
      const int t0 = (tile - row_local * a.tiles_per_row) * kTileM;
      const int t = t0 + m;
      const int slot = edge_slot(t, a.nb_map, a.n_frames);
      mbar_wait(bar(BAR_SCOUT_FULL + pb), (it >> 1) & 1u, a.error_flag, 6);
      if (warp == kProducerWarp0 && lane == 0) FE_TRACE(12, it, 0);
      float bound = 0.0f;
      if (slot >= 0) {
        bound = s_side_max[pb * 4 + slot];
      } else if (t < a.n_frames) {
        const int c_first = max(0, (t0 - 1) * hop) / kCellFloats;
        const int ca = ((t - 1) * hop) / kCellFloats - c_first, cb = ((t + 1) * hop - 1) / kCellFloats - c_first;
        for (int cidx = ca; cidx <= cb; ++cidx) bound = fmaxf(bound, s_cells[pb * kMaxCells + cidx]);
      }
      float scale, unscale;
      fe_gemm_frame_scale(2.0f * bound, scale, unscale);
      // edge frames read their (reflect-padded) samples from the side buffer, laid out like one box row
      const unsigned char* side_f = reinterpret_cast<const unsigned char*>(s_side + (pb * side_slots + (slot < 0 ? 0 : slot)) * side_floats);
      const int my_swz = slot < 0 ? swz : 0;
      float mid_re = 0.0f, mid_im = 0.0f, carry = 0.0f;
#pragma unroll 1
      for (int q = 0; q < a.nstages; ++q, ++n) {
        const uint32_t s = n & 1u, par = (n >> 1) & 1u;
        mbar_wait(bar(BAR_SAMP_FULL + s), par, a.error_flag, 7);
        if (warp == kProducerWarp0 && lane == 0) FE_TRACE(1, it, q);
        const unsigned char* fbox = slot < 0 ? smem + L.samp + s * kSampStageBytes + m * 128 : side_f + q * 128;
        const unsigned char* bbox = slot < 0 ? fbox + kSampBoxBytes : side_f + a.kpairs * 4 + q * 128;
        // forward elements 16*half .. +15 = x[c + j0 + i].  Backward: x[c - j0 - i] is box element
        // 32 - 16*half - i: elements 16*(1-half) .. +15 give i = 1..15 (reversed); i = 0 is element 16 of the
        // box (second half) or the first element of the previous stage's box / the centre sample (first half)
        float fwd[16], bwd[16], buf[16];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float4 f = *reinterpret_cast<const float4*>(fbox + (((4 * half + ch) ^ my_swz) << 4));
          fwd[4 * ch + 0] = f.x; fwd[4 * ch + 1] = f.y; fwd[4 * ch + 2] = f.z; fwd[4 * ch + 3] = f.w;
          const float4 bq = *reinterpret_cast<const float4*>(bbox + (((4 * (1 - half) + ch) ^ my_swz) << 4));
          buf[4 * ch + 0] = bq.x; buf[4 * ch + 1] = bq.y; buf[4 * ch + 2] = bq.z; buf[4 * ch + 3] = bq.w;
        }
        if (half == 0) {
          bwd[0] = (q == 0) ? fwd[0] : carry;
          carry = *reinterpret_cast<const float*>(bbox + ((0 ^ my_swz) << 4));   // element 0: x[c - 32(q+1)]
        } else {
          bwd[0] = *reinterpret_cast<const float*>(bbox + ((4 ^ my_swz) << 4));  // element 16
        }
#pragma unroll
        for (int i = 1; i < 16; ++i) bwd[i] = buf[16 - i];
        // samples are in registers: release the sample slot to the loader
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_SAMP_EMPTY + s));
        fe_u4 chunk[8];
        fe_gemm_produce_half(fwd, bwd, scale, 32 * q + 16 * half, mid_re_w, mid_im_w, mid_re, mid_im, chunk);
        // the A slot is free once the MMAs of its previous use have retired
        mbar_wait(bar(BAR_STAGE_EMPTY + s), par ^ 1u, a.error_flag, 10);
        unsigned char* a_row = smem + L.a_stage + s * kAStageBytes + half * kTileM * 16 + m * 16;
#pragma unroll
        for (int sf = 0; sf < 8; ++sf) *reinterpret_cast<fe_u4*>(a_row + sf * fe_gemm_tile_bytes(kTileM)) = chunk[sf];
        if (q == a.nstages - 1) {
          if (half == 0) s_unscale[pb * pe_stride + m] = unscale * unscale;
          s_p128[(pb * pe_stride) * 2 + half * kTileM + m] = make_float2(mid_re, mid_im);
        }
        fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (warp == kProducerWarp0 && lane == 0) FE_TRACE(2, it, q);
        if (lane == 0) {
          mbar_arrive(bar(BAR_A_FULL + s));
          // cells were consumed at the top of the tile, the side buffers are read at every stage
          if (q == a.nstages - 1) mbar_arrive(bar(BAR_SCOUT_EMPTY + pb));
        }
      }
    }
  } else {
    // ================================ epilogue (warps 0..7) =========================================
    const int ew = warp - kEpilogueWarp0;
    const int quarter = warp & 3;          // TMEM lanes 32*quarter .. +31 are the ones this warp may read
    const int grp = ew >> 2;               // column half
    const int m = quarter * 32 + lane;
    const int kper = a.nhalf / 2;
    const int k_begin = grp * kper, k_end = k_begin + kper;
    const int etid = ew * 32 + lane;
    const int nfil = a.n_filter;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const uint32_t pb = it & 1u;
      const int row_local = tile / a.tiles_per_row;
      const int t0 = (tile - row_local * a.tiles_per_row) * kTileM;
      const int valid_rows = min(kTileM, a.n_frames - t0);
      mbar_wait(bar(BAR_ACC_FULL), it & 1u, a.error_flag, 8);
      tc_fence_after();
      if (ew == 0 && lane == 0) FE_TRACE(4, it, 0);
      const float us2 = s_unscale[pb * pe_stride + m];
      const float2 pm0 = s_p128[(pb * pe_stride) * 2 + m], pm1 = s_p128[(pb * pe_stride) * 2 + kTileM + m];
      const float mid_r = pm0.x + pm1.x, mid_i = pm0.y + pm1.y;
      const float p_mid = fmaf(mid_r, mid_r, mid_i * mid_i);
      // each column group sums into its own [filter][frame] array (the two groups meet on the filters
      // around their boundary); a thread only ever touches its own frame's column, so plain updates do
      float* ecol = s_energy + grp * nfil * kTileM + m;
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const fe_gemm_fbw* fbw = a.fb_in_smem ? s_fb : g_fb;
      float alo[FE_GEMM_FB_SPAN], ahi[FE_GEMM_FB_SPAN];
#pragma unroll 1
      for (int k0 = k_begin; k0 < k_end; k0 += 8) {
        float ce[8], co[8], se[8], so[8];
        tmem_ld8(tbase + (uint32_t)(0 * a.nhalf + k0), ce);
        tmem_ld8(tbase + (uint32_t)(1 * a.nhalf + k0), co);
        tmem_ld8(tbase + (uint32_t)(2 * a.nhalf + k0), se);
        tmem_ld8(tbase + (uint32_t)(3 * a.nhalf + k0), so);
        if ((k0 & (FE_GEMM_CHUNK - 1)) == 0) {
#pragma unroll
          for (int j = 0; j < FE_GEMM_FB_SPAN; ++j) alo[j] = ahi[j] = 0.0f;
        }
        tmem_ld_wait();
        fe_gemm_epi_cols<8>(fbw + k0, ce, co, se, so, alo, ahi);
        if ((k0 & (FE_GEMM_CHUNK - 1)) == FE_GEMM_CHUNK - 8) {
          // end of a 16-column chunk: add its 4 + 4 sums to the frame's filter sums
          const int c = k0 / FE_GEMM_CHUNK;
          const int bl = s_ctl->base_lo[c], bh = s_ctl->base_hi[c];
#pragma unroll
          for (int j = 0; j < FE_GEMM_FB_SPAN; ++j) {
            if (bl + j < nfil) ecol[(bl + j) * kTileM] += alo[j] * us2;
            if (bh + j < nfil) ecol[(bh + j) * kTileM] += ahi[j] * us2;
          }
        }
      }
      // accumulators are in registers now: hand TMEM back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (ew == 0 && lane == 0) FE_TRACE(5, it, 0);
      if (lane == 0) mbar_arrive(bar(BAR_ACC_EMPTY));
      if (grp == 1) {
        // bin n_fft/4, evaluated by the producer in true units
#pragma unroll
        for (int j = 0; j < FE_GEMM_FB_SPAN; ++j)
          if (s_ctl->mid_base + j < nfil) ecol[(s_ctl->mid_base + j) * kTileM] += p_mid * s_ctl->mid_w[j];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // all emissions of the tile are in s_energy
      // coalesced store: consecutive threads -> consecutive frames of one filter
      float* dst = a.energies + ((size_t)row_local * nfil) * a.n_frames + t0;
      float vmax = 0.0f;
      for (int i = etid; i < nfil * kTileM; i += kNumEpilogueWarps * 32) {
        const int f = i >> 7, r = i & (kTileM - 1);
        const float v = s_energy[i] + s_energy[nfil * kTileM + i];
        s_energy[i] = 0.0f;
        s_energy[nfil * kTileM + i] = 0.0f;
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
      if (ew == 0 && lane == 0) FE_TRACE(6, it, 0);
    }
  }

  // ---- teardown -------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
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
  // measured on B200 (profiles/): energies kernel 2.8 M utt/s (dft_gemm) vs 1.1 M utt/s (fft) on the LFCC
  // configuration, so AUTO takes the tensor-core variant wherever it is supported
  return true;
}
bool fe_gemm_variant_built(void) { return true; }
bool fe_gemm_auto_prefers(const b200fe_params* p) { return fe_gemm_supported(p) && fe_gemm_preferred(p); }

int64_t fe_gemm_workspace_bytes(const b200fe_params* p, int64_t chunk_rows, int64_t T) {
  (void)p; (void)chunk_rows; (void)T;
  return 65536;  // error flag (+ trace buffer in FE_GEMM_TRACE builds)
}

cudaError_t fe_gemm_launch(const b200fe_params* p, const fe_fft_args& fa, int64_t row_base, int64_t rows,
                           void* gemm_ws, cudaStream_t stream, int* launches) {
  *launches = 0;
  encode_tiled_fn enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  const int hop = p->hop_length;
  const int64_t T = fa.T;
  const int nb_map = (int)((T - 1) / hop);  // hop blocks b whose column `hop` (= first sample of block b+1) exists
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
  a.nb_map = nb_map;

  // 3-D tensor map over the chunk's rows: [hop samples][nb_map hop blocks][rows]; box 32 x 128 x 1, swizzle 128B
  CUtensorMap map;
  const cuuint64_t gdim[3] = {(cuuint64_t)hop, (cuuint64_t)nb_map, (cuuint64_t)rows};
  const cuuint64_t gstride[2] = {(cuuint64_t)hop * 4, (cuuint64_t)T * 4};
  const cuuint32_t box[3] = {32, (cuuint32_t)kTileM, 1};
  const cuuint32_t estride[3] = {1, 1, 1};
  void* gbase = (void*)(fa.wave + row_base * T);
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, gbase, gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;

  a.fb_in_smem = make_layout(a.nhalf, a.kpairs, a.n_filter, 1 + a.n_frames - a.nb_map, true).total <= 227 * 1024;
  const smem_layout L = make_layout(a.nhalf, a.kpairs, a.n_filter, 1 + a.n_frames - a.nb_map, a.fb_in_smem != 0);
  const int smem = L.total;
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
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
#ifdef FE_GEMM_TRACE
  cudaMemsetAsync(a.error_flag, 0, 65536, stream);
#endif
  if (e != cudaSuccess) return e;
  // the kernel addresses wave relative to the chunk: rows are local to the tensor map
  gemm_args b = a;
  b.wave = fa.wave;  // absolute pointer + (row_base + row_local) * T in the kernel's direct-load paths
  fe_gemm_kernel<<<grid, kThreads, smem, stream>>>(map, b);
  *launches = 1;
  return cudaGetLastError();
}
