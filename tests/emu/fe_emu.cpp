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
