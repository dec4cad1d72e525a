// CPU emulation of fe_fft_kernel<1> + fe_tail_kernel for tests (TEST INFRASTRUCTURE, not product).
//
// It includes the very same phase functions the CUDA kernels are made of
// (csrc/fe_fft.cuh, csrc/fe_tail.cuh) and runs them with lanes / threads as plain loops and
// __syncwarp / __syncthreads as the boundaries between those loops.  This checks the index logic of
// the kernels (Stockham addressing, real-FFT split, band filterbank, replicate-clamped delta tiles,
// reflect / repeat-pad staging) on a machine without a GPU.  Built by tests/conftest.py with g++.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "fe_fft.cuh"
#include "fe_tail.cuh"

extern "C" int fe_emu_features(const float* wave, int64_t R, int64_t T, const int64_t* offsets,
                               const int32_t* lengths, const b200fe_params* p, const void* tables,
                               int ft, int tt, float* power_out /* [R][n_freq][nF] or NULL */,
                               float* out /* [R][n_out][nF] */) {
  const unsigned char* blob = (const unsigned char*)tables;
  const fe_blob_header* h = (const fe_blob_header*)blob;
  if (h->magic != FE_BLOB_MAGIC) return -1;
  const int n_fft = p->n_fft, nh = n_fft / 2, n_freq = nh + 1, hop = p->hop_length;
  const int nF = (int)(1 + T / hop);
  const int nfil = p->n_filter, ncoef = p->n_coef;
  const int nc = ncoef > 0 ? ncoef : nfil;
  const int n_out = nc * (1 + p->deltas);
  const float* s_win = (const float*)(blob + h->off_window);
  const fe_c2* s_tw = (const fe_c2*)(blob + h->off_twiddle);
  const fe_c2* s_rtw = (const fe_c2*)(blob + h->off_rtwiddle);
  const int32_t* bstart = (const int32_t*)(blob + h->off_band_start);
  const int32_t* blen = (const int32_t*)(blob + h->off_band_len);
  const int32_t* bwoff = (const int32_t*)(blob + h->off_band_woff);
  const float* bw = (const float*)(blob + h->off_band_w);
  const float* dct = (const float*)(blob + h->off_dct);
  const bool radix2_first = (__builtin_ctz(nh) & 1) != 0;
  const int nthreads = 256, nwarps = 8;
  const bool need_max = p->log_mode == B200FE_LOG_DB && p->top_db >= 0.0f;
  const int64_t ngroups = (R + p->top_db_group - 1) / p->top_db_group;

  std::vector<float> energies((size_t)R * nfil * nF);
  std::vector<float> gmax(ngroups, 0.0f);

  // ---- fe_fft_kernel<1> ------------------------------------------------------------------------
  const int tiles = (nF + ft - 1) / ft;
  std::vector<float> s_stage((size_t)(ft - 1) * hop + n_fft);
  std::vector<fe_c2> b0(nh + 1), b1(nh + 1);
  std::vector<float> s_tile((size_t)(nfil > n_freq ? nfil : n_freq) * (ft + 1));
  for (int64_t row = 0; row < R; ++row) {
    const float* src = offsets ? wave + offsets[row] : wave + row * T;
    const int clip_len = offsets ? lengths[row] : (int)T;
    for (int tile = 0; tile < tiles; ++tile) {
      const int t0 = tile * ft;
      const int nf_here = ft < nF - t0 ? ft : nF - t0;
      const int seg_here = (nf_here - 1) * hop + n_fft;
      for (int tid = 0; tid < nthreads; ++tid)
        fe_stage_load(tid, nthreads, src, clip_len, (int)T, n_fft, t0 * hop, seg_here, p->preemph, s_stage.data());
      for (int warp = 0; warp < nwarps; ++warp) {
        for (int fl = warp; fl < nf_here; fl += nwarps) {
          const float* frame = s_stage.data() + (size_t)fl * hop;
          fe_c2* in = b0.data();
          fe_c2* o = b1.data();
          for (int lane = 0; lane < 32; ++lane) fe_fft_stage_first(lane, frame, s_win, o, nh, radix2_first);
          int ns = radix2_first ? 2 : 4;
          while (ns < nh) {
            fe_c2* t = in; in = o; o = t;
            for (int lane = 0; lane < 32; ++lane) fe_fft_stage4(lane, in, o, s_tw, nh, ns);
            ns <<= 2;
          }
          float* pw = (float*)in;
          for (int lane = 0; lane < 32; ++lane) fe_fft_power(lane, o, s_rtw, pw, nh);
          if (power_out)
            for (int k = 0; k < n_freq; ++k) power_out[((size_t)row * n_freq + k) * nF + t0 + fl] = pw[k];
          for (int lane = 0; lane < 32; ++lane)
            fe_fbank_apply(lane, pw, bstart, blen, bwoff, bw, nfil, s_tile.data() + fl, ft + 1);
        }
      }
      for (int c = 0; c < nfil; ++c)
        for (int t = 0; t < nf_here; ++t) {
          const float v = s_tile[(size_t)c * (ft + 1) + t];
          energies[((size_t)row * nfil + c) * nF + t0 + t] = v;
          float& g = gmax[row / p->top_db_group];
          g = v > g ? v : g;
        }
    }
  }

  // ---- fe_tail_kernel --------------------------------------------------------------------------
  const int n = p->deltas > 0 ? (p->delta_win - 1) / 2 : 1;
  const int halo = p->deltas * n;
  const int w = tt + 2 * halo;
  const int tthreads = 128;
  std::vector<float> s_e((size_t)nfil * w), s_cbuf((size_t)nc * w), s_d((size_t)nc * w);
  const int ttiles = (nF + tt - 1) / tt;
  for (int64_t row = 0; row < R; ++row) {
    float floor_db = -INFINITY;
    if (need_max) floor_db = 10.0f * log10f(fmaxf(gmax[row / p->top_db_group], 1e-10f)) - p->top_db;
    for (int tile = 0; tile < ttiles; ++tile) {
      const int t0 = tile * tt, tv0 = t0 - halo;
      const float* src = energies.data() + (size_t)row * nfil * nF;
      for (int tid = 0; tid < tthreads; ++tid)
        fe_tail_load(tid, tthreads, src, nfil, nF, w, tv0, p->log_mode, floor_db, s_e.data());
      float* s_c = s_e.data();
      if (ncoef > 0) {
        s_c = s_cbuf.data();
        for (int tid = 0; tid < tthreads; ++tid) fe_tail_dct(tid, tthreads, s_e.data(), dct, nfil, ncoef, w, s_c);
      }
      const int nt_here = tt < nF - t0 ? tt : nF - t0;
      if (p->deltas >= 1)
        for (int tid = 0; tid < tthreads; ++tid) fe_tail_delta(tid, tthreads, s_c, nc, w, n, tv0, nF, s_d.data());
      float* out_row = out + (size_t)row * n_out * nF;
      for (int tid = 0; tid < tthreads; ++tid)
        fe_tail_store(tid, tthreads, s_c, s_d.data(), nc, w, n, halo, t0, nt_here, nF, p->deltas, out_row);
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// CPU emulation of the DFT-GEMM variant (fe_stream.cu): the same tile geometry and per-thread functions
// (fe_gemm.cuh: stream tiles over the flat frame sequence, hop-block slots, per-frame scale, production units,
// sliding even/odd filterbank drain) and the same operand images / tables (fe_gemm_layout.h), with tcgen05.mma
// replaced by loops that decode the fp16 operand tiles at their UMMA layout offsets and accumulate in fp32.
// Output: filterbank energies [R][n_filter][nF] (what fe_stream_kernel writes to the workspace).
// ---------------------------------------------------------------------------------------------------
#include "fe_gemm.cuh"

static float emu_half_at(const unsigned char* img, int off) {
  __half h;
  memcpy(&h, img + off, 2);
  return __half2float(h);
}

extern "C" int fe_emu_gemm_energies(const float* wave, int64_t R, int64_t T_, const b200fe_params* p,
                                    const void* tables, float* energies) {
  const unsigned char* blob = (const unsigned char*)tables;
  const fe_blob_header* h = (const fe_blob_header*)blob;
  if (h->magic != FE_BLOB_MAGIC || !h->gemm_ok) return -1;
  const int T = (int)T_, hop = p->hop_length, nF = 1 + T / hop, nfil = p->n_filter;
  const int kpairs = h->gemm_kpairs, nhalf = h->gemm_nhalf, nstages = kpairs / 32;
  const fe_drain_w* dw = (const fe_drain_w*)(blob + h->off_gemm_dw);
  const uint32_t* dctl = (const uint32_t*)(blob + h->off_gemm_dctl);
  const fe_drain_hdr* hdr = (const fe_drain_hdr*)(blob + h->off_gemm_dids);
  const float* dwn = (const float*)(blob + h->off_gemm_dwn);
  const float* gmid = (const float*)(blob + h->off_gemm_mid);
  const unsigned char* gB = blob + h->off_gemm_b;
  const int M = FE_GEMM_TILE_M;
  std::vector<float> midc(kpairs);   // interleaved weights of bin n_fft/4, as the kernel builds them in shared memory
  for (int j = 0; j < kpairs; ++j) midc[j] = (j & 1) ? gmid[kpairs + j] : gmid[j];

  const int total = (int)(R * nF), tf = fe_tile_frames(nF);
  const int n_tiles = (total + tf - 1) / tf;
  std::vector<unsigned char> a_stage(fe_gemm_a_stage_bytes());
  std::vector<float> D((size_t)4 * M * nhalf);
  const int kProducerGroups = 2;   // fe_stream.cu: production unit (stage, khalf) belongs to warp group khalf
  std::vector<float> samp, bmax;
  for (int tile = 0; tile < n_tiles; ++tile) {
    const fe_tile_geo g = fe_tile_geometry(tile, tf, total, nF);
    if (g.nv > 132) return -2;
    // ---- loader: the tile's hop blocks, one slot each (edge blocks reflect-padded)
    samp.assign((size_t)g.nv * hop, 0.0f);
    bmax.assign(g.nv, 0.0f);
    for (int s = 0; s < g.nv; ++s) {
      const int sv = g.sv0 + s, row = sv / (nF + 1), v = sv - row * (nF + 1);
      const float* x = wave + (int64_t)row * T_;
      for (int e = 0; e < hop; ++e) {
        int idx = (v - 1) * hop + e;
        idx = idx < 0 ? -idx : idx;
        idx = idx >= T ? 2 * (T - 1) - idx : idx;
        const float val = x[idx];
        samp[(size_t)s * hop + e] = val;
        bmax[s] = fmaxf(bmax[s], fabsf(val));   // scout
      }
    }
    std::fill(D.begin(), D.end(), 0.0f);
    std::vector<float> scale(M), unscale(M), mre((size_t)kProducerGroups * M, 0.0f), mim((size_t)kProducerGroups * M, 0.0f);
    std::vector<int> slot(M);
    for (int m = 0; m < M; ++m) {
      const int mm = m < g.count ? m : g.count - 1;
      const int row = (g.g0 + mm) / nF;
      slot[m] = mm + (row - g.row0);
      fe_gemm_frame_scale(2.0f * fmaxf(bmax[slot[m]], bmax[slot[m] + 1]), scale[m], unscale[m]);
    }
    for (int q = 0; q < nstages; ++q) {
      // producers: thread (m, khalf)
      for (int m = 0; m < M; ++m) {
        const float* brow = samp.data() + (size_t)slot[m] * hop;
        const float* frow = brow + hop;
        for (int khalf = 0; khalf < 2; ++khalf) {
          const int j0 = 32 * q + 16 * khalf;
          float fwd[16], bwd[16];
          for (int i = 0; i < 16; ++i) fwd[i] = frow[j0 + i];
          bwd[0] = (j0 == 0) ? fwd[0] : brow[hop - j0];
          for (int i = 1; i < 16; ++i) bwd[i] = brow[hop - j0 - i];
          fe_u4 chunk[8];
          const int grp = khalf;   // each group accumulates its own partial of bin n_fft/4
          fe_stream_produce_unit(fwd, bwd, scale[m], midc.data() + j0, mre[(size_t)grp * M + m], mim[(size_t)grp * M + m], chunk);
          for (int sf = 0; sf < 8; ++sf)
            memcpy(a_stage.data() + sf * fe_gemm_tile_bytes(M) + fe_gemm_operand_offset(M, m, 8 * khalf), &chunk[sf], 16);
        }
      }
      // "tcgen05.mma": D_sub += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo over the stage's 16 K values
      const unsigned char* b_stage = gB + (size_t)q * fe_gemm_b_stage_bytes(nhalf);
      for (int sub = 0; sub < 4; ++sub)
        for (int m = 0; m < M; ++m)
          for (int n = 0; n < nhalf; ++n) {
            float acc = D[((size_t)sub * M + m) * nhalf + n];
            for (int kk = 0; kk < 16; ++kk) {
              const float ah = emu_half_at(a_stage.data(), fe_gemm_a_tile_offset(sub, 0) + fe_gemm_operand_offset(M, m, kk));
              const float al = emu_half_at(a_stage.data(), fe_gemm_a_tile_offset(sub, 1) + fe_gemm_operand_offset(M, m, kk));
              const float bh = emu_half_at(b_stage, fe_gemm_b_tile_offset(nhalf, sub, 0) + fe_gemm_operand_offset(nhalf, n, kk));
              const float bl = emu_half_at(b_stage, fe_gemm_b_tile_offset(nhalf, sub, 1) + fe_gemm_operand_offset(nhalf, n, kk));
              acc += ah * bh + al * bh + ah * bl;
            }
            D[((size_t)sub * M + m) * nhalf + n] = acc;
          }
    }
    // drain: thread (frame m, run); finished segments are the filters' final energies
    for (int m = 0; m < g.count; ++m) {
      const float us2 = unscale[m] * unscale[m];
      const int gi = g.g0 + m, row = gi / nF, t = gi - row * nF;
      auto emit = [&](int f, float v) {
        if (f < nfil) energies[((size_t)row * nfil + f) * nF + t] = v * us2;
      };
      float left[2][2];
      const int nbatch = nhalf / FE_DRAIN_BATCH, npairs = nhalf / 2;
      for (int run = 0; run < 2; ++run) {
        float l0[2] = {0.0f, 0.0f};
        for (int half = 0; half < 2; ++half) {
          fe_drain_state st;
          fe_drain_init(st, *hdr, run, half);
          for (int b = half * nbatch / 2; b < (half + 1) * nbatch / 2; ++b) {
            fe_f2 pw[4];
            for (int p = 0; p < 4; ++p) {
              const int k = b * FE_DRAIN_BATCH + 2 * p;
              const fe_f2 c0 = fe_f2{D[((size_t)0 * M + m) * nhalf + k], D[((size_t)0 * M + m) * nhalf + k + 1]};
              const fe_f2 c1 = fe_f2{D[((size_t)1 * M + m) * nhalf + k], D[((size_t)1 * M + m) * nhalf + k + 1]};
              const fe_f2 s0 = fe_f2{D[((size_t)2 * M + m) * nhalf + k], D[((size_t)2 * M + m) * nhalf + k + 1]};
              const fe_f2 s1 = fe_f2{D[((size_t)3 * M + m) * nhalf + k], D[((size_t)3 * M + m) * nhalf + k + 1]};
              pw[p] = run == 0 ? fe_drain_power<0>(c0, c1, s0, s1) : fe_drain_power<1>(c0, c1, s0, s1);
            }
            fe_drain_batch(pw, dw + run * npairs + b * 4, dwn + (size_t)(run * nbatch + b) * 16, dctl[run * nbatch + b], st, emit);
          }
          if (half == 0) {
            for (int c = 0; c < 2; ++c) l0[c] = fe_drain_leftover(st, c, 0.0f, 0.0f);
          } else {
            float p_mid = 0.0f;
            if (run == 0) {
              const float bs = (float)(1 << FE_GEMM_B_SCALE_LOG2);
              const float re = (mre[m] + mre[(size_t)M + m]) * bs;
              const float im = (mim[m] + mim[(size_t)M + m]) * bs;
              p_mid = fmaf(re, re, im * im);
            }
            fe_drain_join_halves(st, *hdr, run, l0, p_mid, left[run], emit);
          }
        }
      }
      for (int c = 0; c < 2; ++c) {
        if (hdr->merge[c]) {
          emit(hdr->last[0][c], left[0][c] + left[1][c]);
        } else {
          emit(hdr->last[0][c], left[0][c]);
          emit(hdr->last[1][c], left[1][c]);
        }
      }
    }
  }
  return 0;
}
