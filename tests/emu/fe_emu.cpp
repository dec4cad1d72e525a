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
// CPU emulation of the DFT-GEMM variant (fe_gemm.cu): the same per-thread functions (fe_gemm.cuh) and the
// same operand images / tables (fe_gemm_layout.h), with tcgen05.mma replaced by loops that decode the
// fp16 operand tiles at their UMMA layout offsets and accumulate in fp32.  Output: filterbank energies
// [R][n_filter][nF] (what fe_gemm_kernel writes to the workspace).
// ---------------------------------------------------------------------------------------------------
#include "fe_gemm.cuh"

static float emu_half_at(const unsigned char* img, int off) {
  __half h;
  memcpy(&h, img + off, 2);
  return __half2float(h);
}

static int emu_reflect(int s, int T) {
  if (s < 0) s = -s;
  if (s >= T) s = 2 * (T - 1) - s;
  return s;
}

static float emu_absmax(const float* x, int lo, int hi) {
  float m = 0.0f;
  for (int i = lo; i < hi; ++i) m = fmaxf(m, fabsf(x[i]));
  return m;
}

extern "C" int fe_emu_gemm_energies(const float* wave, int64_t R, int64_t T_, const b200fe_params* p,
                                    const void* tables, float* energies) {
  const unsigned char* blob = (const unsigned char*)tables;
  const fe_blob_header* h = (const fe_blob_header*)blob;
  if (h->magic != FE_BLOB_MAGIC || !h->gemm_ok) return -1;
  const int T = (int)T_, hop = p->hop_length, nF = 1 + T / hop, nfil = p->n_filter;
  const int kpairs = h->gemm_kpairs, nhalf = h->gemm_nhalf, nstages = kpairs / 32;
  const int nb_full = T / hop;
  const fe_gemm_fbw* fbw = (const fe_gemm_fbw*)(blob + h->off_gemm_fb);
  const fe_gemm_fbctl* ctl = (const fe_gemm_fbctl*)(blob + h->off_gemm_fbflag);
  const float* mid = (const float*)(blob + h->off_gemm_mid);
  const unsigned char* gB = blob + h->off_gemm_b;
  const int M = FE_GEMM_TILE_M;
  std::vector<unsigned char> a_stage(fe_gemm_a_stage_bytes());
  std::vector<float> D((size_t)4 * M * nhalf), E((size_t)FE_GEMM_MAX_FILTERS * M);
  const int tiles = (nF + M - 1) / M;
  for (int64_t row = 0; row < R; ++row) {
    const float* x = wave + row * T_;
    for (int tile = 0; tile < tiles; ++tile) {
      const int t0 = tile * M;
      std::fill(D.begin(), D.end(), 0.0f);
      std::fill(E.begin(), E.end(), 0.0f);
      std::vector<float> scale(M), unscale(M), mre(M, 0.0f), mim(M, 0.0f);
      // scout + frame scale
      std::vector<float> bm(M + 1);
      for (int s = 0; s <= M; ++s) {
        const int b = t0 - 1 + s;
        if (b < 0) bm[s] = emu_absmax(x, 0, std::min(T, 2 * hop + 1));
        else if (b >= nb_full) bm[s] = emu_absmax(x, std::max(0, (nb_full - 2) * hop), T);
        else bm[s] = emu_absmax(x, b * hop, (b + 1) * hop);
      }
      for (int m = 0; m < M; ++m) fe_gemm_frame_scale(2.0f * fmaxf(bm[m], bm[m + 1]), scale[m], unscale[m]);
      for (int q = 0; q < nstages; ++q) {
        // producers
        for (int m = 0; m < M; ++m) {
          const int t = t0 + m, c = t * hop;
          const bool valid = t < nF;
          for (int half = 0; half < 2; ++half) {
            const int j0 = 32 * q + 16 * half;
            float fwd[16], bwd[16];
            for (int i = 0; i < 16; ++i) {
              fwd[i] = valid ? x[emu_reflect(c + j0 + i, T)] : 0.0f;
              bwd[i] = valid ? x[emu_reflect(c - j0 - i, T)] : 0.0f;
            }
            fe_u4 chunk[8];
            fe_gemm_produce_half(fwd, bwd, scale[m], j0, mid, mid + kpairs, mre[m], mim[m], chunk);
            for (int sf = 0; sf < 8; ++sf)
              memcpy(a_stage.data() + sf * fe_gemm_tile_bytes(M) + fe_gemm_operand_offset(M, m, 8 * half), &chunk[sf], 16);
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
      // epilogue: per frame, chunks of 16 columns
      for (int m = 0; m < M; ++m) {
        const float us2 = unscale[m] * unscale[m];
        for (int c = 0; c < nhalf / FE_GEMM_CHUNK; ++c) {
          float ce[16], co[16], se[16], so[16], alo[FE_GEMM_FB_SPAN], ahi[FE_GEMM_FB_SPAN];
          for (int i = 0; i < 16; ++i) {
            const int k = 16 * c + i;
            ce[i] = D[((size_t)0 * M + m) * nhalf + k];
            co[i] = D[((size_t)1 * M + m) * nhalf + k];
            se[i] = D[((size_t)2 * M + m) * nhalf + k];
            so[i] = D[((size_t)3 * M + m) * nhalf + k];
          }
          for (int j = 0; j < FE_GEMM_FB_SPAN; ++j) alo[j] = ahi[j] = 0.0f;
          fe_gemm_epi_cols<8>(fbw + 16 * c, ce, co, se, so, alo, ahi);
          fe_gemm_epi_cols<8>(fbw + 16 * c + 8, ce + 8, co + 8, se + 8, so + 8, alo, ahi);
          for (int j = 0; j < FE_GEMM_FB_SPAN; ++j) {
            if (ctl->base_lo[c] + j < nfil) E[(size_t)(ctl->base_lo[c] + j) * M + m] += alo[j] * us2;
            if (ctl->base_hi[c] + j < nfil) E[(size_t)(ctl->base_hi[c] + j) * M + m] += ahi[j] * us2;
          }
        }
        const float pmid = mre[m] * mre[m] + mim[m] * mim[m];
        for (int j = 0; j < FE_GEMM_FB_SPAN; ++j)
          if (ctl->mid_base + j < nfil) E[(size_t)(ctl->mid_base + j) * M + m] += pmid * ctl->mid_w[j];
      }
      const int valid_rows = std::min(M, nF - t0);
      for (int f = 0; f < nfil; ++f)
        for (int r = 0; r < valid_rows; ++r) energies[((size_t)row * nfil + f) * nF + t0 + r] = E[(size_t)f * M + r];
    }
  }
  return 0;
}
