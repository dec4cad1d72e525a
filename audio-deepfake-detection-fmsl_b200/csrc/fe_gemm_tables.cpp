// Host-side construction of the DFT-GEMM variant's constant tables (layout documented in
// fe_gemm_layout.h).  Pure host code: runs and is tested without a GPU.
#include <math.h>
#include <string.h>
#include <vector>

#include <cuda_fp16.h>

#include "fe_gemm_layout.h"
#include "fe_gemm_tables.h"

namespace {

struct gemm_geom {
  bool ok = false;
  int kpairs = 0;   // folded sample pairs = win_length / 2
  int nhalf = 0;    // n_fft / 4
  int nstages = 0;  // kpairs / 32
};

gemm_geom geometry(const b200fe_params* p) {
  gemm_geom g;
  if (p->n_filter < 1 || p->n_filter > FE_GEMM_MAX_FILTERS) return g;
  if (p->win_length != 2 * p->hop_length) return g;  // frame = exactly two hop blocks (staging contract)
  if (p->win_length > p->n_fft) return g;
  const int kpairs = p->win_length / 2;
  const int nhalf = p->n_fft / 4;
  if (kpairs % 32 != 0 || kpairs < 32 || kpairs > 256) return g;
  if (nhalf % 16 != 0 || nhalf < 32 || nhalf > 128) return g;  // 4 accumulators of nhalf columns fit TMEM
  if (p->preemph != 0.0f) return g;
  g.ok = true;
  g.kpairs = kpairs;
  g.nhalf = nhalf;
  g.nstages = kpairs / 32;
  return g;
}

}  // namespace

int64_t fe_gemm_plan_layout(const b200fe_params* p, fe_blob_header* h, int64_t off) {
  h->gemm_ok = 0;
  const gemm_geom g = geometry(p);
  if (!g.ok) return off;
  h->gemm_kpairs = g.kpairs;
  h->gemm_nhalf = g.nhalf;
  off = (off + 127) & ~(int64_t)127;
  h->off_gemm_b = (int32_t)off;
  h->gemm_b_bytes = g.nstages * fe_gemm_b_stage_bytes(g.nhalf);
  off = fe_align16(off + h->gemm_b_bytes);
  h->off_gemm_fb = (int32_t)off;
  off = fe_align16(off + (int64_t)(g.nhalf + 1) * sizeof(fe_gemm_fb_entry));
  h->off_gemm_mid = (int32_t)off;
  off = fe_align16(off + (int64_t)2 * g.kpairs * 4);
  h->gemm_ok = 1;  // provisional: fe_gemm_pack clears it when the window / filterbank do not qualify
  return off;
}

int32_t fe_gemm_pack(const b200fe_params* p, fe_blob_header* h, const float* window, const float* fbank,
                     char* base) {
  if (!h->gemm_ok) return B200FE_OK;
  const gemm_geom g = geometry(p);
  const int n_fft = p->n_fft, n_freq = n_fft / 2 + 1, nfil = p->n_filter;
  // ---- symmetric window about the frame centre -------------------------------------------------
  // centred window index i = j + win/2 for offset j from the centre; j = -win/2 is the lone sample
  std::vector<double> wj(g.kpairs);
  const int half = p->win_length / 2;
  double wmax = 0;
  for (int i = 0; i < p->win_length; ++i) wmax = fmax(wmax, fabs((double)window[i]));
  if (fabs((double)window[0]) > 1e-7 * wmax) { h->gemm_ok = 0; return B200FE_OK; }  // lone sample must vanish
  wj[0] = window[half];
  for (int j = 1; j < g.kpairs; ++j) {
    const double a = window[half + j], b = window[half - j];
    if (fabs(a - b) > 1e-6 * wmax) { h->gemm_ok = 0; return B200FE_OK; }
    wj[j] = 0.5 * (a + b);
  }
  // ---- filterbank structure: every bin feeds at most two adjacent filters, monotonically ---------
  std::vector<int> phi(n_freq);
  int cur = -1;
  for (int b = 0; b < n_freq; ++b) {
    int first = -1, last = -1, cnt = 0;
    for (int f = 0; f < nfil; ++f)
      if (fbank[(int64_t)b * nfil + f] != 0.0f) { if (first < 0) first = f; last = f; ++cnt; }
    if (cnt == 0) { phi[b] = cur; continue; }
    if (last - first > 1 || cnt > 2) { h->gemm_ok = 0; return B200FE_OK; }
    int want;
    if (cnt == 2) want = first;
    else want = (cur < first - 1) ? first - 1 : (cur > first ? -2 : cur);  // keep cur if it still covers `first`
    if (want == -2 || want < cur) { h->gemm_ok = 0; return B200FE_OK; }
    cur = want;
    phi[b] = cur;
  }
  auto fbw = [&](int b, int f) -> float { return (f >= 0 && f < nfil) ? fbank[(int64_t)b * nfil + f] : 0.0f; };
  fe_gemm_fb_entry* fb = (fe_gemm_fb_entry*)(base + h->off_gemm_fb);
  const int nyq = n_fft / 2;
  for (int k = 0; k <= g.nhalf; ++k) {
    const int bl = k, bh = nyq - k;
    fb[k].w_lo_a = fbw(bl, phi[bl]);
    fb[k].w_lo_b = fbw(bl, phi[bl] + 1);
    fb[k].w_hi_a = fbw(bh, phi[bh]);
    fb[k].w_hi_b = fbw(bh, phi[bh] + 1);
    fb[k].phi_lo = phi[bl];
    fb[k].phi_hi = phi[bh];
    const int adv_lo = k > 0 ? phi[bl] - phi[bl - 1] : 0;
    const int adv_hi = k > 0 ? phi[bh + 1] - phi[bh] : 0;
    // the kernel's sweep moves a window by at most one filter per bin
    if (adv_lo < 0 || adv_lo > 1 || adv_hi < 0 || adv_hi > 1) { h->gemm_ok = 0; return B200FE_OK; }
    fb[k].adv = adv_lo | (adv_hi << 1);
    fb[k].pad0 = 0;
  }
  // ---- bin n_fft/4 (handled on the CUDA cores): true-unit weights ------------------------------
  float* mid = (float*)(base + h->off_gemm_mid);
  for (int j = 0; j < g.kpairs; ++j) {
    const double hj = (j == 0) ? 0.5 : 1.0;
    mid[j] = (float)(hj * wj[j] * cos(M_PI * 0.5 * j));             // Re X[n/4] = sum a_e[j] * mid[j]
    mid[g.kpairs + j] = (float)(-wj[j] * sin(M_PI * 0.5 * j));      // Im X[n/4] = sum a_o[j] * mid[kp + j]
  }
  // ---- DFT operand tiles: fp16 hi/lo of 2^14 * window * cos/sin, UMMA K-major no-swizzle ----------
  char* bt = base + h->off_gemm_b;
  const double scale = ldexp(1.0, FE_GEMM_B_SCALE_LOG2);
  for (int q = 0; q < g.nstages; ++q) {
    for (int sub = 0; sub < 4; ++sub) {
      const bool is_sin = sub >= 2, odd = (sub & 1) != 0;
      for (int kk = 0; kk < 16; ++kk) {
        const int j = 32 * q + 2 * kk + (odd ? 1 : 0);
        for (int n = 0; n < g.nhalf; ++n) {
          const double th = 2.0 * M_PI * (double)((int64_t)n * j % n_fft) / (double)n_fft;
          double v;
          if (!is_sin) v = ((j == 0) ? 0.5 : 1.0) * wj[j] * cos(th);
          else v = -wj[j] * sin(th);
          v *= scale;
          const __half hi = __float2half_rn((float)v);
          const __half lo = __float2half_rn((float)(v - (double)__half2float(hi)));
          const int64_t o_hi = (int64_t)q * fe_gemm_b_stage_bytes(g.nhalf) + fe_gemm_b_tile_offset(g.nhalf, sub, 0) +
                               fe_gemm_operand_offset(g.nhalf, n, kk);
          const int64_t o_lo = (int64_t)q * fe_gemm_b_stage_bytes(g.nhalf) + fe_gemm_b_tile_offset(g.nhalf, sub, 1) +
                               fe_gemm_operand_offset(g.nhalf, n, kk);
          memcpy(bt + o_hi, &hi, 2);
          memcpy(bt + o_lo, &lo, 2);
        }
      }
    }
  }
  return B200FE_OK;
}
