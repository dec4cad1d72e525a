// Streaming tcgen05 kernel of the DFT-GEMM variant (sm_100a): waveform -> filterbank energies.
//
// The launch's rows are one stream of frames (row-major: utterance, frame); a tile = 128 consecutive stream
// frames = the 128 TMEM lanes of one M=128 MMA, whatever utterances they belong to, so every tile is full.
// One persistent CTA per SM walks tiles  blockIdx.x, blockIdx.x + gridDim.x, ...
//
//   loader warp   per tile: the hop blocks the tile's frames need, once each, as 1-D bulk copies (TMA) into
//                 padded shared-memory rows (conflict-free lane <-> frame reads); reflect-padded edge blocks
//                 are synthesised with plain loads.  Per stage: the 32 KB of DFT operand tiles.
//   MMA warp      one thread: 12 tcgen05.mma (M=128, N=n_fft/4, K=16; 4 sub-GEMMs x 3 split-fp16 products) per
//                 stage into the 4 TMEM accumulators, tcgen05.commit -> mbarriers
//   16 worker warps, all doing the same thing in phases:
//     scout       max|x| per group of 4 hop blocks (shared memory) -> per-frame power-of-two scale
//     produce     fold + scale + fp16 hi/lo split of 16 sample pairs per thread into the UMMA A tiles; the two
//                 halves of the warps (0-7 / 8-15) take alternate stages, i.e. alternate A slots
//     drain       tcgen05.ld of the four accumulators, powers, sliding even/odd triangular-filter sums
//                 (fe_gemm_layout.h), all 16 warps: warp = (TMEM lane quarter, column group)
//     finalize    energies of the tile -> workspace [row][filter][frame] (+ per-group maximum for top_db)
// TMEM holds exactly the four accumulators (4 x 128 columns), so the MMAs of a tile and its drain cannot
// overlap; everything else does: production runs one stage ahead of the MMAs, the next tile's samples and
// operand stages are loaded during MMA tail and drain.
#include "fe_tc.cuh"

#include "fe_gemm.cuh"
#include "fe_gemm.h"
#include "fe_gemm_tables.h"

namespace {

constexpr int kWorkerWarps = 16;
constexpr int kWorkerThreads = kWorkerWarps * 32;
constexpr int kMmaWarp = 16, kLoaderWarp = 17;
constexpr int kThreads = 18 * 32;
constexpr int kTileM = FE_GEMM_TILE_M;
constexpr int kMaxSlots = 132;                    // hop blocks of a tile: 128 + 1 + one more per utterance boundary
constexpr int kSlotGroups = kMaxSlots / 4;        // scout granularity: 4 hop blocks
constexpr int kMinFrames = 43;                    // a tile then spans at most 4 utterances
constexpr int kAStageBytes = 8 * 2 * kTileM * 16; // 32 KB: [sub 4][hi, lo] tiles of 128 rows x 16 K

struct stream_args {
  const float* wave;        // first row of the launch
  const void* tables;
  float* energies;          // [rows][n_filter][n_frames]
  unsigned int* group_max;  // or NULL
  int* error_flag;
  int64_t T;
  int64_t row_base;         // absolute index of the launch's first row (group_max indexing)
  int32_t rows, n_frames, n_filter, hop, nhalf, nstages, kpairs;
  int32_t total_frames, n_tiles, top_db_group;
};

struct smem_layout {
  int samp, a_stage, b_stage, dw, dids, dctl, mid, gmax, us2, midp, bars, tmem_slot, total;
};

__host__ __device__ inline smem_layout make_layout(int hop, int nhalf, int kpairs) {
  smem_layout L;
  int off = 0;
  L.samp = off;      off += kMaxSlots * (hop * 4 + 16);
  off = (off + 127) & ~127;
  L.a_stage = off;   off += 2 * kAStageBytes;
  L.b_stage = off;   off += 2 * fe_gemm_b_stage_bytes(nhalf);
  L.dw = off;        off += (nhalf + 1) * (int)sizeof(fe_drain_w);
  L.dids = off;      off += ((nhalf + 1) * (int)sizeof(fe_drain_ids) + 15) & ~15;
  L.dctl = off;      off += ((nhalf / 8 + 1) * 4 + 15) & ~15;
  L.mid = off;       off += 2 * kpairs * 4;
  L.gmax = off;      off += ((kSlotGroups + 1) * 4 + 15) & ~15;
  L.us2 = off;       off += kTileM * 4;
  L.midp = off;      off += 4 * kTileM * 8;   // [producer half-group 2][K half 2][frame] (Re, Im) partials of bin n_fft/4
  L.bars = off;      off += 16 * 8;
  L.tmem_slot = off; off += 16;
  L.total = off;
  return L;
}

enum { BAR_SAMP_FULL = 0, BAR_SAMP_EMPTY = 1, BAR_A_FULL = 2, BAR_B_FULL = 4, BAR_STAGE_FREE = 6, BAR_ACC_FULL = 8,
       BAR_ACC_EMPTY = 9, BAR_COUNT = 10 };

#ifdef FE_GEMM_TRACE
#define ST_TRACE(ev, it, q) do { if (blockIdx.x == 0 && (it) < 8) { ((long long*)(a.error_flag + 64))[((it) * 8 + (q)) * 16 + (ev)] = clock64(); } } while (0)
#else
#define ST_TRACE(ev, it, q) do { } while (0)
#endif

struct tile_geo {
  int g0, count, row0, row_last, sv0, nv;
};
__device__ __forceinline__ tile_geo tile_geometry(int tile, int total_frames, int nF) {
  tile_geo t;
  t.g0 = tile * kTileM;
  t.count = min(kTileM, total_frames - t.g0);
  t.row0 = t.g0 / nF;
  const int g_last = t.g0 + t.count - 1;
  t.row_last = g_last / nF;
  // hop block v of row r (samples [(v-1) hop, v hop)) has stream index r (nF+1) + v; frame (r, t) reads blocks t, t+1
  t.sv0 = t.g0 + t.row0;
  t.nv = t.count + (t.row_last - t.row0) + 1;
  return t;
}

__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkerThreads) : "memory"); }

// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) fe_stream_kernel(const stream_args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const smem_layout L = make_layout(a.hop, a.nhalf, a.kpairs);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned char* blob = reinterpret_cast<const unsigned char*>(a.tables);
  const fe_blob_header* h = reinterpret_cast<const fe_blob_header*>(blob);
  const int nF = a.n_frames, hop = a.hop, nfil = a.n_filter;
  const int rs = hop * 4 + 16;   // bytes per hop-block row: +16 puts 8 consecutive rows on 8 distinct 16-byte bank groups

  if (!h->stream_ok) {   // tables without the drain tables: report instead of computing garbage
    if (tid == 0) atomicExch(a.error_flag, 98);
    return;
  }
  unsigned char* s_samp = smem + L.samp;
  fe_drain_w* s_dw = reinterpret_cast<fe_drain_w*>(smem + L.dw);
  fe_drain_ids* s_dids = reinterpret_cast<fe_drain_ids*>(smem + L.dids);
  uint32_t* s_dctl = reinterpret_cast<uint32_t*>(smem + L.dctl);
  float* s_mid = reinterpret_cast<float*>(smem + L.mid);
  float* s_gmax = reinterpret_cast<float*>(smem + L.gmax);
  float* s_us2 = reinterpret_cast<float*>(smem + L.us2);
  float2* s_midp = reinterpret_cast<float2*>(smem + L.midp);
  float* s_E = reinterpret_cast<float*>(smem + L.a_stage);   // drain scratch [2][n_filter][128] aliases A slot 0
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L.tmem_slot);
  const uint32_t bars = smem_u32(smem + L.bars);
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

  // ---- one-time setup -----------------------------------------------------------------------------
  {
    const fe_drain_w* gdw = reinterpret_cast<const fe_drain_w*>(blob + h->off_gemm_dw);
    for (int i = tid; i <= a.nhalf; i += kThreads) s_dw[i] = gdw[i];
    const fe_drain_ids* gids = reinterpret_cast<const fe_drain_ids*>(blob + h->off_gemm_dids);
    for (int i = tid; i <= a.nhalf; i += kThreads) s_dids[i] = gids[i];
    const uint32_t* gctl = reinterpret_cast<const uint32_t*>(blob + h->off_gemm_dctl);
    for (int i = tid; i <= a.nhalf / 8; i += kThreads) s_dctl[i] = gctl[i];
    const float* gmid = reinterpret_cast<const float*>(blob + h->off_gemm_mid);
    for (int i = tid; i < 2 * a.kpairs; i += kThreads) s_mid[i] = gmid[i];
    // the scout reads whole groups of 4 rows including the row pads and rows a tile does not use: keep them finite
    float4* z = reinterpret_cast<float4*>(s_samp);
    for (int i = tid; i < kMaxSlots * rs / 16; i += kThreads) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();
  }
  if (tid == 0) {
    mbar_init(bar(BAR_SAMP_FULL), 2);
    mbar_init(bar(BAR_SAMP_EMPTY), kWorkerWarps);
    mbar_init(bar(BAR_A_FULL + 0), kWorkerWarps / 2);
    mbar_init(bar(BAR_A_FULL + 1), kWorkerWarps / 2);
    mbar_init(bar(BAR_B_FULL + 0), 1);
    mbar_init(bar(BAR_B_FULL + 1), 1);
    mbar_init(bar(BAR_STAGE_FREE + 0), 1);
    mbar_init(bar(BAR_STAGE_FREE + 1), 1);
    mbar_init(bar(BAR_ACC_FULL), 1);
    mbar_init(bar(BAR_ACC_EMPTY), 1);
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
  const uint32_t b_stage_bytes = (uint32_t)fe_gemm_b_stage_bytes(a.nhalf);

  if (warp == kLoaderWarp) {
    // ================================ loader ==========================================================
    const unsigned char* gB = blob + h->off_gemm_b;
    const uint32_t row_bytes = (uint32_t)hop * 4u;
    uint32_t n = 0, it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const tile_geo g = tile_geometry(tile, a.total_frames, nF);
      mbar_wait(bar(BAR_SAMP_EMPTY), (it & 1u) ^ 1u, a.error_flag, 1);   // workers are done with the previous tile's samples
      ST_TRACE(0, it, 0);
      int n_edge = 0;
      for (int row = g.row0; row <= g.row_last; ++row) {
        const int sa = row * (nF + 1) - g.sv0, sb = sa + nF;
        n_edge += (sa >= 0 && sa < g.nv) + (sb >= 0 && sb < g.nv);
      }
      fence_proxy_async();   // earlier generic writes of edge rows vs. the bulk copies below
      if (lane == 0) mbar_arrive_expect_tx(bar(BAR_SAMP_FULL), (uint32_t)(g.nv - n_edge) * row_bytes);
      __syncwarp();
      for (int s = lane; s < g.nv; s += 32) {
        const int sv = g.sv0 + s;
        const int row = sv / (nF + 1), v = sv - row * (nF + 1);
        if (v >= 1 && v < nF)
          bulk_g2s(smem_u32(s_samp + s * rs), a.wave + (int64_t)row * a.T + (int64_t)(v - 1) * hop, row_bytes,
                   bar(BAR_SAMP_FULL));
      }
      // edge blocks: v = 0 (reflect about sample 0) and v = nF (tail of the utterance + reflect about sample T-1)
      for (int row = g.row0; row <= g.row_last; ++row) {
        const float* x = a.wave + (int64_t)row * a.T;
        for (int e2 = 0; e2 < 2; ++e2) {
          const int v = e2 ? nF : 0;
          const int s = row * (nF + 1) + v - g.sv0;
          if (s < 0 || s >= g.nv) continue;
          float* dst = reinterpret_cast<float*>(s_samp + s * rs);
          for (int e = lane; e < hop; e += 32) {
            int idx = (v - 1) * hop + e;
            idx = idx < 0 ? -idx : idx;
            idx = idx >= T ? 2 * (T - 1) - idx : idx;
            dst[e] = __ldg(x + idx);
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(BAR_SAMP_FULL));
        for (int q = 0; q < a.nstages; ++q, ++n) {
          const uint32_t s = n & 1u, par = (n >> 1) & 1u;
          mbar_wait(bar(BAR_STAGE_FREE + s), par ^ 1u, a.error_flag, 2);   // the MMAs that read slot s have retired
          mbar_arrive_expect_tx(bar(BAR_B_FULL + s), b_stage_bytes);
          bulk_g2s(smem_u32(smem + L.b_stage + s * b_stage_bytes), gB + (size_t)q * b_stage_bytes, b_stage_bytes,
                   bar(BAR_B_FULL + s));
        }
      }
      __syncwarp();
    }
  } else if (warp == kMmaWarp) {
    // ================================ MMA issuer ======================================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(a.nhalf >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t b_lbo = (uint32_t)a.nhalf * 16u;
      uint32_t n = 0, it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(bar(BAR_ACC_EMPTY), (it & 1u) ^ 1u, a.error_flag, 3);   // previous tile drained
        tc_fence_after();
        for (int q = 0; q < a.nstages; ++q, ++n) {
          const uint32_t s = n & 1u, par = (n >> 1) & 1u;
          mbar_wait(bar(BAR_B_FULL + s), par, a.error_flag, 4);
          mbar_wait(bar(BAR_A_FULL + s), par, a.error_flag, 5);
          tc_fence_after();
          ST_TRACE(3, it, q);
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
          umma_commit(bar(BAR_STAGE_FREE + s));
          if (q == a.nstages - 1) umma_commit(bar(BAR_ACC_FULL));
        }
      }
    }
  } else {
    // ================================ workers (warps 0..15) ===========================================
    const int quarter = warp & 3;          // TMEM lane quarter (drain) = frame quarter (production)
    const int khalf = (warp >> 2) & 1;     // production: which 16 of the stage's 32 sample pairs
    const int pgrp = warp >> 3;            // production: stages with (n & 1) == pgrp, i.e. A slot pgrp
    const int cg = warp >> 2;              // drain: column group
    const int m = quarter * 32 + lane;
    const int cpg = a.nhalf / FE_DRAIN_GROUPS;
    uint32_t n = 0, it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const tile_geo g = tile_geometry(tile, a.total_frames, nF);
      const int mm = min(m, g.count - 1);                 // rows past the end of the stream repeat the last frame
      const int row = (g.g0 + mm) / nF;
      const int slot = mm + (row - g.row0);               // backward hop block; the forward one is slot + 1
      mbar_wait(bar(BAR_SAMP_FULL), it & 1u, a.error_flag, 6);
      if (tid == 0) ST_TRACE(1, it, 0);
      // ---- scout: max |x| per group of 4 rows (pads and unused rows hold zeros / stale finite samples)
      for (int grp = warp; grp * 4 < g.nv + 1; grp += kWorkerWarps) {
        const float4* p = reinterpret_cast<const float4*>(s_samp + grp * 4 * rs);
        const int n4 = rs / 4;   // float4 per 4 rows
        float mx = 0.0f;
        for (int i = lane; i < n4; i += 32) {
          const float4 v = p[i];
          mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
        mx = warp_max(mx);
        if (lane == 0) s_gmax[grp] = mx;
      }
      worker_bar();
      if (tid == 0) ST_TRACE(7, it, 0);
      float scale, unscale;
      fe_gemm_frame_scale(2.0f * fmaxf(s_gmax[slot >> 2], s_gmax[(slot + 1) >> 2]), scale, unscale);
      // ---- produce
      const unsigned char* brow = s_samp + slot * rs;
      const unsigned char* frow = brow + rs;
      float mid_re = 0.0f, mid_im = 0.0f;
#pragma unroll 1
      for (int q = 0; q < a.nstages; ++q, ++n) {
        if ((int)(n & 1u) != pgrp) continue;
        const uint32_t par = (n >> 1) & 1u;
        const int j0 = 32 * q + 16 * khalf;
        float fwd[16], bwd[16], buf[16];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float4 f = *reinterpret_cast<const float4*>(frow + j0 * 4 + ch * 16);
          fwd[4 * ch + 0] = f.x; fwd[4 * ch + 1] = f.y; fwd[4 * ch + 2] = f.z; fwd[4 * ch + 3] = f.w;
          const float4 b = *reinterpret_cast<const float4*>(brow + (hop - j0 - 16) * 4 + ch * 16);
          buf[4 * ch + 0] = b.x; buf[4 * ch + 1] = b.y; buf[4 * ch + 2] = b.z; buf[4 * ch + 3] = b.w;
        }
        // bwd[i] = x[c - j0 - i] = backward-row element hop - j0 - i; element hop (i = 0, j0 = 0) is the centre sample
        bwd[0] = (j0 == 0) ? fwd[0] : *reinterpret_cast<const float*>(brow + (hop - j0) * 4);
#pragma unroll
        for (int i = 1; i < 16; ++i) bwd[i] = buf[16 - i];
        fe_u4 chunk[8];
        fe_stream_produce_unit(fwd, bwd, scale, s_mid + j0, s_mid + a.kpairs + j0, mid_re, mid_im, chunk);
        mbar_wait(bar(BAR_STAGE_FREE + pgrp), par ^ 1u, a.error_flag, 7);   // MMAs of this slot's previous use retired
        unsigned char* a_row = smem + L.a_stage + pgrp * kAStageBytes + khalf * kTileM * 16 + m * 16;
#pragma unroll
        for (int sf = 0; sf < 8; ++sf) *reinterpret_cast<fe_u4*>(a_row + sf * fe_gemm_tile_bytes(kTileM)) = chunk[sf];
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_A_FULL + pgrp));
        if (tid == 0 || tid == 256) ST_TRACE(2, it, q);
      }
      s_midp[(pgrp * 2 + khalf) * kTileM + m] = make_float2(mid_re, mid_im);
      if (pgrp == 0 && khalf == 0) s_us2[m] = unscale * unscale;
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(BAR_SAMP_EMPTY));
      // ---- drain
      mbar_wait(bar(BAR_ACC_FULL), it & 1u, a.error_flag, 8);
      tc_fence_after();
      if (tid == 0) ST_TRACE(4, it, 0);
      {
        float4* z = reinterpret_cast<float4*>(s_E);
        for (int i = tid; i < 2 * nfil * kTileM / 4; i += kWorkerThreads) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      worker_bar();
      if (tid == 0) ST_TRACE(8, it, 0);
      {
        fe_drain_state st;
#pragma unroll
        for (int j = 0; j < 4; ++j) { st.acc[j] = 0.0f; st.id[j] = -1; }
        float* e_col = s_E + (cg & 1) * nfil * kTileM + m;
        const float us2 = s_us2[m];
        const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int k_begin = cg * cpg;
#pragma unroll 1
        for (int k0 = k_begin; k0 < k_begin + cpg; k0 += 8) {
          float ce[8], co[8], se[8], so[8];
          tmem_ld8(tbase + (uint32_t)(0 * a.nhalf + k0), ce);
          tmem_ld8(tbase + (uint32_t)(1 * a.nhalf + k0), co);
          tmem_ld8(tbase + (uint32_t)(2 * a.nhalf + k0), se);
          tmem_ld8(tbase + (uint32_t)(3 * a.nhalf + k0), so);
          const unsigned ctl = s_dctl[k0 >> 3];
          tmem_ld_wait();
          fe_drain_cols<8>(s_dw + k0, s_dids + k0, ctl, ce, co, se, so, st, e_col, nfil, us2);
        }
        if (cg == FE_DRAIN_GROUPS - 1) {
          // bin n_fft/4 from the producers' partial sums (scaled sample units -> accumulator units: x 2^14)
          float re = 0.0f, im = 0.0f;
#pragma unroll
          for (int p = 0; p < 4; ++p) { const float2 v = s_midp[p * kTileM + m]; re += v.x; im += v.y; }
          const float bs = (float)(1 << FE_GEMM_B_SCALE_LOG2);
          re *= bs; im *= bs;
          fe_drain_mid(s_dw + a.nhalf, s_dids + a.nhalf, s_dctl[a.nhalf >> 3], fmaf(re, re, im * im), st, e_col, nfil, us2);
        }
        fe_drain_flush(st, e_col, nfil, us2);
      }
      tc_fence_before();
      worker_bar();
      if (tid == 0) {
        mbar_arrive(bar(BAR_ACC_EMPTY));   // TMEM is free for the next tile's MMAs
        ST_TRACE(5, it, 0);
      }
      // ---- finalize: this tile's energies -> workspace, per-group maximum
      {
        const int t = (g.g0 + mm) - row * nF;
        const bool valid = m < g.count;
        float* dst = a.energies + (size_t)row * nfil * nF + t;
        float vmax = 0.0f;
        for (int f = warp >> 2; f < nfil; f += kWorkerWarps / 4) {   // thread -> frame m, filters f, f+4, ...
          const float v = s_E[f * kTileM + m] + s_E[(nfil + f) * kTileM + m];
          if (valid) {
            dst[(size_t)f * nF] = v;
            vmax = fmaxf(vmax, v);
          }
        }
        if (a.group_max) {
          const int grp_id = (int)((a.row_base + row) / a.top_db_group);
          const int grp0 = __shfl_sync(0xffffffffu, grp_id, 0);
          if (__all_sync(0xffffffffu, grp_id == grp0)) {
            vmax = warp_max(vmax);
            if (lane == 0) atomicMax(a.group_max + grp0, __float_as_uint(vmax));
          } else if (valid) {
            atomicMax(a.group_max + grp_id, __float_as_uint(vmax));
          }
        }
      }
      if (tid == 0) ST_TRACE(6, it, 0);
      // (the next tile's scout barrier separates these reads of s_E from the next production's A stores)
    }
  }

  // ---- teardown -------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace

bool fe_stream_supported(const b200fe_params* p, int64_t T, int64_t rows) {
  if (!fe_gemm_supported(p)) return false;
  const int nhalf = p->n_fft / 4, kpairs = p->win_length / 2;
  if (nhalf % (8 * FE_DRAIN_GROUPS) != 0) return false;
  const int64_t nF = 1 + T / p->hop_length;
  if (nF < kMinFrames || rows * nF >= (int64_t)1 << 30) return false;
  if (T <= p->hop_length || (T & 3) != 0) return false;
  if (2 * p->n_filter * kTileM * 4 > kAStageBytes) return false;
  return make_layout(p->hop_length, nhalf, kpairs).total <= 227 * 1024;
}

cudaError_t fe_stream_launch(const b200fe_params* p, const fe_fft_args& fa, int64_t row_base, int64_t rows,
                             void* gemm_ws, cudaStream_t stream, int* launches) {
  *launches = 0;
  stream_args a;
  a.wave = fa.wave + row_base * fa.T;
  a.tables = fa.tables;
  a.energies = fa.out;
  a.group_max = fa.group_max;
  a.error_flag = (int*)gemm_ws;
  a.T = fa.T;
  a.row_base = row_base;
  a.rows = (int32_t)rows;
  a.n_frames = fa.n_frames;
  a.n_filter = p->n_filter;
  a.hop = p->hop_length;
  a.nhalf = p->n_fft / 4;
  a.kpairs = p->win_length / 2;
  a.nstages = a.kpairs / 32;
  a.total_frames = (int32_t)(rows * fa.n_frames);
  a.n_tiles = (a.total_frames + kTileM - 1) / kTileM;
  a.top_db_group = fa.top_db_group;

  const int smem = make_layout(a.hop, a.nhalf, a.kpairs).total;
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  static int attr_done = 0;
  if (attr_done < smem) {
    cudaError_t e = cudaFuncSetAttribute(fe_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done = smem;
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.n_tiles < sms ? a.n_tiles : sms;
#ifdef FE_GEMM_TRACE
  cudaError_t e = cudaMemsetAsync(a.error_flag, 0, 65536, stream);
#else
  cudaError_t e = cudaMemsetAsync(a.error_flag, 0, 4, stream);
#endif
  if (e != cudaSuccess) return e;
  fe_stream_kernel<<<grid, kThreads, smem, stream>>>(a);
  *launches = 1;
  return cudaGetLastError();
}
