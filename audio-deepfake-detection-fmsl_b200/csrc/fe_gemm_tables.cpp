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
  if (nhalf % (8 * FE_DRAIN_GROUPS) != 0 || nhalf < 32 || nhalf > 128) return g;  // 4 accumulators of nhalf columns fit TMEM
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
  h->off_gemm_mid = (int32_t)off;
  off = fe_align16(off + (int64_t)2 * g.kpairs * 4);
  h->off_gemm_dw = (int32_t)off;
  off = fe_align16(off + (int64_t)(g.nhalf / 2 + 1) * sizeof(fe_drain_w));
  h->off_gemm_dctl = (int32_t)off;
  off = fe_align16(off + (int64_t)(g.nhalf / 8 + 1) * 4);
  h->off_gemm_dids = (int32_t)off;
  off = fe_align16(off + (int64_t)(g.nhalf / 2 + 1) * sizeof(fe_drain_ids));
  h->gemm_ok = 1;  // provisional: fe_gemm_pack clears it when the window / filterbank do not qualify
  return off;
}

// Drain tables (fe_gemm_layout.h): per column pair the weights of the four accumulator classes x two halves, the
// switch flags and the filter rows after the switches.  Returns false when the filterbank does not qualify (a bin
// with two filters of the same parity, or emission buffers that do not fit the A slots they alias).
static bool pack_drain_tables(fe_blob_header* h, const float* fbank, int nfil, int nhalf, char* base) {
  if (nhalf % (8 * FE_DRAIN_GROUPS) != 0 || nfil > FE_GEMM_MAX_FILTERS) return false;
  fe_drain_w* dw = (fe_drain_w*)(base + h->off_gemm_dw);
  uint32_t* dctl = (uint32_t*)(base + h->off_gemm_dctl);
  fe_drain_ids* dids = (fe_drain_ids*)(base + h->off_gemm_dids);
  const int npairs = nhalf / 2;
  memset(dw, 0, (size_t)(npairs + 1) * sizeof(fe_drain_w));
  memset(dctl, 0, (size_t)(nhalf / 8 + 1) * 4);
  const int nyq = 2 * nhalf, ppg = npairs / FE_DRAIN_GROUPS;   // pairs per column group
  std::vector<unsigned> touched(nfil, 0u);
  // filter of parity `par` with weight on `bin` (-1: none, -2: more than one)
  auto filter_of = [&](int bin, int par) {
    int f_found = -1;
    for (int f = par; f < nfil; f += 2)
      if (fbank[(int64_t)bin * nfil + f] != 0.0f) { if (f_found >= 0) return -2; f_found = f; }
    return f_found;
  };
  for (int g = 0; g < FE_DRAIN_GROUPS; ++g) {
    int cur[4][2] = {{-1, -1}, {-1, -1}, {-1, -1}, {-1, -1}};
    // the last group also takes bin n_fft/4 as the even column of one more (half) pair
    const int p_end = (g + 1) * ppg + (g == FE_DRAIN_GROUPS - 1 ? 1 : 0);
    for (int p = g * ppg; p < p_end; ++p) {
      unsigned flags = 0;
      for (int hh = 0; hh < 2; ++hh) {
        const int k = 2 * p + hh;
        if (k > nhalf) continue;                     // the odd half of the bin-n_fft/4 entry does not exist
        for (int run = 0; run < 2; ++run) {
          if (k == nhalf && run == 1) continue;      // bin n_fft/4 belongs to the ascending run only
          const int bin = run == 0 ? k : nyq - k;
          for (int par = 0; par < 2; ++par) {
            const int a = 2 * run + par;
            const int f = filter_of(bin, par);
            if (f == -2) return false;
            if (f < 0) continue;
            dw[p].w[a][hh] = fbank[(int64_t)bin * nfil + f];
            touched[f] |= 1u << g;
            if (f != cur[a][hh]) {
              if (p != g * ppg) flags |= 1u << (2 * a + hh);   // the group's first pair: the walk starts aimed at it
              cur[a][hh] = f;
            }
          }
        }
      }
      for (int a = 0; a < 4; ++a)
        for (int hh = 0; hh < 2; ++hh)
          dids[p].off[2 * a + hh] = (int16_t)((cur[a][hh] < 0 ? nfil : cur[a][hh]) * FE_GEMM_TILE_M * 4);
      dctl[p / 4] |= flags << (8 * (p % 4));
    }
  }
  // Column groups that share a filter must emit into different buffers.  Two buffers indexed by group parity do
  // when a filter is only ever shared by adjacent groups; otherwise every group gets its own buffer (if the
  // n_filter x 128 arrays still fit the 64 KB of A slots they alias).
  bool adjacent_only = true;
  for (int f = 0; f < nfil; ++f) {
    const unsigned m = touched[f];
    if (m == 0) continue;
    const unsigned low = m & (0u - m);
    if (m != low && m != (low | (low << 1))) adjacent_only = false;
  }
  h->gemm_nbuf = adjacent_only ? 2 : FE_DRAIN_GROUPS;
  return h->gemm_nbuf * (nfil + 1) * FE_GEMM_TILE_M * 4 <= 2 * 8 * fe_gemm_tile_bytes(FE_GEMM_TILE_M);   // both A slots (+1: the dummy row)
}

int32_t fe_gemm_pack(const b200fe_params* p, fe_blob_header* h, const float* window, const float* fbank,
                     char* base) {
  if (!h->gemm_ok) return B200FE_OK;
  const gemm_geom g = geometry(p);
  const int n_fft = p->n_fft, nfil = p->n_filter;
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
  // ---- filterbank: sliding even/odd accumulator tables of the drain ------------------------------
  if (!pack_drain_tables(h, fbank, nfil, g.nhalf, base)) { h->gemm_ok = 0; return B200FE_OK; }
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
