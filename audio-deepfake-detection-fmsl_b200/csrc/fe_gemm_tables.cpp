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
  if (nhalf % 16 != 0 || nhalf < 32 || nhalf > 128) return g;  // 4 accumulators of nhalf columns fit TMEM; two halves of whole batches
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
  off = fe_align16(off + (int64_t)2 * (g.nhalf / 2) * sizeof(fe_drain_w));
  h->off_gemm_dctl = (int32_t)off;
  off = fe_align16(off + (int64_t)2 * (g.nhalf / FE_DRAIN_BATCH) * 4);
  h->off_gemm_dids = (int32_t)off;
  off = fe_align16(off + (int64_t)sizeof(fe_drain_hdr));
  h->off_gemm_dwn = (int32_t)off;
  off = fe_align16(off + (int64_t)2 * (g.nhalf / FE_DRAIN_BATCH) * 16 * 4);
  h->gemm_ok = 1;  // provisional: fe_gemm_pack clears it when the window / filterbank do not qualify
  return off;
}

// Drain tables (fe_gemm_layout.h): per run the pair weights of the two filter-parity classes, per batch the boundary
// flags, new targets and the weights behind the boundary.  Returns false when the filterbank does not qualify: a bin
// with two filters of the same parity, a filter in two separate ranges of a run, two boundaries of one class in one
// batch window, or a filter without any bin (its energy would never be stored).
static bool pack_drain_tables(fe_blob_header* h, const float* fbank, int nfil, int nhalf, char* base) {
  if (nhalf % (2 * FE_DRAIN_BATCH) != 0 || nfil > FE_GEMM_MAX_FILTERS || nfil >= FE_DRAIN_NONE) return false;
  const int npairs = nhalf / 2, nbatch = nhalf / FE_DRAIN_BATCH, nyq = 2 * nhalf;
  fe_drain_w* dw = (fe_drain_w*)(base + h->off_gemm_dw);
  uint32_t* dctl = (uint32_t*)(base + h->off_gemm_dctl);
  fe_drain_hdr* hdr = (fe_drain_hdr*)(base + h->off_gemm_dids);
  float* dwn = (float*)(base + h->off_gemm_dwn);
  memset(dw, 0, (size_t)2 * npairs * sizeof(fe_drain_w));
  memset(dctl, 0, (size_t)2 * nbatch * 4);
  memset(hdr, 0, sizeof(*hdr));
  memset(dwn, 0, (size_t)2 * nbatch * 16 * 4);
  // filter of parity `par` with weight on `bin` (-1: none, -2: more than one)
  auto filter_of = [&](int bin, int par) {
    int f_found = -1;
    for (int f = par; f < nfil; f += 2)
      if (fbank[(int64_t)bin * nfil + f] != 0.0f) { if (f_found >= 0) return -2; f_found = f; }
    return f_found;
  };
  std::vector<int> seen(nfil, 0);
  for (int run = 0; run < 2; ++run) {
    for (int par = 0; par < 2; ++par) {
      // columns 0 .. nhalf-1, and for run 0 column nhalf = bin n_fft/4 (added by the thread after its last batch)
      const int ncol = run == 0 ? nhalf + 1 : nhalf;
      std::vector<int> f(ncol, -1);
      for (int k = 0; k < ncol; ++k) {
        f[k] = filter_of(run == 0 ? k : nyq - k, par);
        if (f[k] == -2) return false;
      }
      // filter-less ranges shorter than a batch are absorbed by the following (else the preceding) segment, with zero
      // weights: fewer boundaries
      for (int k = 0; k < ncol;) {
        int e = k;
        while (e + 1 < ncol && f[e + 1] == f[k]) ++e;
        if (f[k] == -1 && e - k + 1 < FE_DRAIN_BATCH) {
          const int repl = e + 1 < ncol ? f[e + 1] : (k > 0 ? f[k - 1] : -1);
          for (int i = k; i <= e; ++i) f[i] = repl;
        }
        k = e + 1;
      }
      // a filter in at most one segment of the class
      std::vector<int> seg_of(nfil, 0);
      for (int k = 0; k < ncol;) {
        int e = k;
        while (e + 1 < ncol && f[e + 1] == f[k]) ++e;
        if (f[k] >= 0) {
          if (seg_of[f[k]]++) return false;
          seen[f[k]] |= 1 << run;
        }
        k = e + 1;
      }
      auto weight = [&](int k, int filt) { return filt < 0 ? 0.0f : fbank[(int64_t)(run == 0 ? k : nyq - k) * nfil + filt]; };
      const int ksplit = nhalf / 2;   // first column of half 1
      hdr->first[run][0][par] = f[0] < 0 ? FE_DRAIN_NONE : f[0];
      hdr->first[run][1][par] = f[ksplit] < 0 ? FE_DRAIN_NONE : f[ksplit];
      hdr->last[run][par] = f[ncol - 1] < 0 ? FE_DRAIN_NONE : f[ncol - 1];
      const bool open = f[ksplit] == f[ksplit - 1] && f[ksplit] >= 0;   // a filter's segment straddles the halves
      hdr->open_tgt[run][par] = open ? f[ksplit] : FE_DRAIN_NONE;
      bool open_pending = open;
      if (run == 0) hdr->wmid[par] = weight(nhalf, f[nhalf]);
      for (int b = 0; b < nbatch; ++b) {
        const int k0 = b * FE_DRAIN_BATCH;
        // boundaries in the window (k0, k0 + 8]: column c is a boundary when f[c] != f[c-1] (c = k0 + 8 may be the
        // run's end: column nhalf of run 0, or nothing for run 1)
        int bcol = -1;
        for (int c = k0 + 1; c <= k0 + FE_DRAIN_BATCH && c < ncol; ++c)
          if (f[c] != f[c - 1]) { if (bcol >= 0) return false; bcol = c; }
        for (int i = 0; i < FE_DRAIN_BATCH; ++i) {
          const int k = k0 + i;
          const bool behind = bcol >= 0 && k >= bcol;
          const float wv = weight(k, f[k]);
          if (!behind) dw[run * npairs + k / 2].w[par][k & 1] = wv;
          else dwn[((run * nbatch + b) * 2 + par) * 8 + i] = wv;   // [pair in batch][half] = column order
        }
        if (bcol >= 0) {
          const uint32_t tgt = f[bcol] < 0 ? FE_DRAIN_NONE : (uint32_t)f[bcol];
          dctl[run * nbatch + b] |= (1u << par) | (tgt << (8 + 8 * par));
          if (k0 >= ksplit && open_pending) {   // half 1's first boundary ends the open segment
            dctl[run * nbatch + b] |= 4u << par;
            open_pending = false;
          }
        }
      }
      hdr->open_last[run][par] = open_pending ? 1 : 0;
    }
  }
  for (int par = 0; par < 2; ++par) {
    const int a = hdr->last[0][par], b = hdr->last[1][par];
    hdr->merge[par] = (a == b && a != FE_DRAIN_NONE) ? 1 : 0;
  }
  // a filter reached by both runs must be the straddler of its class; a filter reached by none is never stored
  for (int f = 0; f < nfil; ++f) {
    if (seen[f] == 0) return false;
    if (seen[f] == 3 && !(hdr->merge[f & 1] && hdr->last[0][f & 1] == f)) return false;
  }
  return true;
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
