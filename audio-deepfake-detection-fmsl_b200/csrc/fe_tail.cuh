// Phases of the feature tail: log / dB + top_db clamp -> DCT-II -> delta -> delta-delta -> store.
// Each function is the share of thread `tid` of `nthreads`; phases are separated by __syncthreads()
// in fe_tail_kernel.  Compiles as plain C++ for the CPU emulation in tests/emu.
//
// Tile geometry: a CTA owns output frames [t0, t0+tt) of one row and works on w = tt + 2*halo tile
// positions; tile position j stands for the "virtual" frame tv0 + j (tv0 = t0 - halo) and holds the
// value at frame clamp(tv0 + j, 0, nF-1) — replicate padding (torchaudio functional.py:1000).
#ifndef FE_TAIL_CUH_
#define FE_TAIL_CUH_

#include <math.h>
#include "fe_common.h"

FE_HD int fe_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// AmplitudeToDB('power'): 10*log10(clamp(x,1e-10)) then max(., floor_db)  (functional.py:356-405);
// log_lf: log(x + 1e-6)  (transforms/_transforms.py:820-822).
FE_HD void fe_tail_load(int tid, int nthreads, const float* src, int nfil, int nF, int w, int tv0,
                        int log_mode, float floor_db, float* s_e) {
  for (int i = tid; i < nfil * w; i += nthreads) {
    const int f = i / w, j = i - f * w;
    const int t = fe_clampi(tv0 + j, 0, nF - 1);
    float v = src[(size_t)f * nF + t];
    if (log_mode == B200FE_LOG_DB) {
      v = 10.0f * log10f(fmaxf(v, 1e-10f));
      v = fmaxf(v, floor_db);
    } else if (log_mode == B200FE_LOG_LN) {
      v = logf(v + 1e-6f);
    }
    s_e[i] = v;
  }
}

// c[k][j] = sum_f e[f][j] * dct[f][k]   (transforms/_transforms.py:827)
FE_HD void fe_tail_dct(int tid, int nthreads, const float* s_e, const float* s_dct, int nfil, int ncoef,
                       int w, float* s_c) {
  for (int i = tid; i < ncoef * w; i += nthreads) {
    const int k = i / w, j = i - k * w;
    float acc = 0.0f;
    for (int f = 0; f < nfil; ++f) acc = fmaf(s_e[(size_t)f * w + j], s_dct[f * ncoef + k], acc);
    s_c[i] = acc;
  }
}

// delta at tile positions [n, w-n): d[j] = sum_m m * c[clamp(clamp(t)+m)] / denom with t = tv0 + j.
// c[clamp(u)] for u = clamp(t)+m sits at tile position u - tv0, which is always inside the tile.
FE_HD void fe_tail_delta(int tid, int nthreads, const float* s_c, int nc, int w, int n, int tv0, int nF,
                         float* s_d) {
  const float denom = (float)(n * (n + 1) * (2 * n + 1)) / 3.0f;
  const int span = w - 2 * n;
  for (int i = tid; i < nc * span; i += nthreads) {
    const int k = i / span, j = n + (i - k * span);
    const int t = fe_clampi(tv0 + j, 0, nF - 1);
    const float* c = s_c + (size_t)k * w + (t - tv0);
    float acc = 0.0f;
    for (int m = -n; m <= n; ++m) acc += (float)m * c[m];
    s_d[(size_t)k * w + j] = acc / denom;
  }
}

// Stores channels [0,nc) = c, [nc,2nc) = delta, [2nc,3nc) = delta-delta for frames t0 .. t0+nt_here.
FE_HD void fe_tail_store(int tid, int nthreads, const float* s_c, const float* s_d, int nc, int w, int n,
                         int halo, int t0, int nt_here, int nF, int deltas, float* out_row) {
  const float denom = (float)(n * (n + 1) * (2 * n + 1)) / 3.0f;
  for (int i = tid; i < nc * nt_here; i += nthreads) {
    const int k = i / nt_here, j = i - k * nt_here;
    out_row[(size_t)k * nF + t0 + j] = s_c[(size_t)k * w + halo + j];
    if (deltas >= 1) out_row[(size_t)(nc + k) * nF + t0 + j] = s_d[(size_t)k * w + halo + j];
    if (deltas >= 2) {
      const float* d = s_d + (size_t)k * w + halo + j;
      float acc = 0.0f;
      for (int m = -n; m <= n; ++m) acc += (float)m * d[m];
      out_row[(size_t)(2 * nc + k) * nF + t0 + j] = acc / denom;
    }
  }
}

#endif  // FE_TAIL_CUH_
