// Host-only part of the C-ABI: parameter validation, error strings and the constant-table packer.
// No CUDA calls here; everything in this file runs (and is tested) on a machine without a GPU.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <vector>

#include "fe_common.h"
#include "fe_gemm_tables.h"

static thread_local char g_err[512] = "";

void fe_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* b200fe_last_error_string(void) { return g_err; }

extern "C" const char* b200fe_status_string(int32_t s) {
  switch (s) {
    case B200FE_OK: return "ok";
    case B200FE_ERR_BAD_ARG: return "bad argument";
    case B200FE_ERR_UNSUPPORTED: return "unsupported configuration (no CPU fallback)";
    case B200FE_ERR_WORKSPACE: return "workspace or tables buffer too small";
    case B200FE_ERR_ALIGNMENT: return "misaligned pointer";
    case B200FE_ERR_CUDA: return "CUDA runtime error";
    case B200FE_ERR_NO_DEVICE: return "no sm_100 device";
    default: return "unknown status";
  }
}

extern "C" int32_t b200fe_version(void) { return B200FE_ABI_VERSION; }

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int32_t fe_validate_params(const b200fe_params* p) {
  if (!p) { fe_set_error("params is NULL"); return B200FE_ERR_BAD_ARG; }
  if (p->abi_version != B200FE_ABI_VERSION) {
    fe_set_error("params.abi_version %d != library %d", p->abi_version, B200FE_ABI_VERSION);
    return B200FE_ERR_BAD_ARG;
  }
  if (!is_pow2(p->n_fft) || p->n_fft < 64 || p->n_fft > 4096) {
    fe_set_error("n_fft=%d: must be a power of two in [64, 4096]", p->n_fft);
    return p->n_fft > 0 ? B200FE_ERR_UNSUPPORTED : B200FE_ERR_BAD_ARG;
  }
  if (p->win_length < 1 || p->win_length > p->n_fft) {
    fe_set_error("win_length=%d must be in [1, n_fft=%d]", p->win_length, p->n_fft);
    return B200FE_ERR_BAD_ARG;
  }
  if (p->hop_length < 1) { fe_set_error("hop_length=%d must be >= 1", p->hop_length); return B200FE_ERR_BAD_ARG; }
  if (p->hop_length > p->n_fft) {
    fe_set_error("hop_length=%d > n_fft=%d is not supported", p->hop_length, p->n_fft);
    return B200FE_ERR_UNSUPPORTED;
  }
  if (p->n_filter < 0 || p->n_filter > 256) {
    fe_set_error("n_filter=%d must be in [0, 256]", p->n_filter);
    return p->n_filter < 0 ? B200FE_ERR_BAD_ARG : B200FE_ERR_UNSUPPORTED;
  }
  if (p->n_coef < 0 || p->n_coef > 256) { fe_set_error("n_coef=%d must be in [0, 256]", p->n_coef); return B200FE_ERR_BAD_ARG; }
  if (p->log_mode < B200FE_LOG_NONE || p->log_mode > B200FE_LOG_LN) {
    fe_set_error("log_mode=%d unknown", p->log_mode);
    return B200FE_ERR_BAD_ARG;
  }
  if (p->deltas < 0 || p->deltas > 2) { fe_set_error("deltas=%d must be 0, 1 or 2", p->deltas); return B200FE_ERR_BAD_ARG; }
  if (p->deltas > 0 && (p->delta_win < 3 || p->delta_win > 9 || (p->delta_win & 1) == 0)) {
    fe_set_error("delta_win=%d must be odd in [3, 9]", p->delta_win);
    return p->delta_win < 3 ? B200FE_ERR_BAD_ARG : B200FE_ERR_UNSUPPORTED;
  }
  if (p->top_db_group < 1) { fe_set_error("top_db_group=%d must be >= 1", p->top_db_group); return B200FE_ERR_BAD_ARG; }
  if (p->variant < B200FE_VARIANT_AUTO || p->variant > B200FE_VARIANT_DFT_GEMM) {
    fe_set_error("variant=%d unknown", p->variant);
    return B200FE_ERR_BAD_ARG;
  }
  if (!(p->preemph == p->preemph)) { fe_set_error("preemph is NaN"); return B200FE_ERR_BAD_ARG; }
  return B200FE_OK;
}

extern "C" int64_t b200fe_n_frames(const b200fe_params* p, int64_t T) {
  int32_t st = fe_validate_params(p);
  if (st != B200FE_OK) return st;
  // reflect padding needs n_fft/2 < T (torch.stft raises otherwise)
  if (T <= p->n_fft / 2) {
    fe_set_error("T=%lld must exceed n_fft/2=%d for reflect padding", (long long)T, p->n_fft / 2);
    return B200FE_ERR_BAD_ARG;
  }
  return 1 + T / p->hop_length;
}

extern "C" int64_t b200fe_n_out_channels(const b200fe_params* p) {
  int32_t st = fe_validate_params(p);
  if (st != B200FE_OK) return st;
  int64_t c = p->n_coef > 0 ? p->n_coef : p->n_filter;
  return c * (1 + p->deltas);
}

// ------------------------------------------------------------------------------------------------
// blob layout
// ------------------------------------------------------------------------------------------------
struct fe_layout {
  fe_blob_header h;
};

static int32_t plan_layout(const b200fe_params* p, const float* fbank, fe_blob_header* h,
                           std::vector<int32_t>* starts, std::vector<int32_t>* lens) {
  memset(h, 0, sizeof(*h));
  const int n_fft = p->n_fft, nh = n_fft / 2, n_freq = nh + 1;
  h->magic = FE_BLOB_MAGIC;
  h->abi_version = B200FE_ABI_VERSION;
  h->n_fft = n_fft;
  h->win_length = p->win_length;
  h->hop_length = p->hop_length;
  h->n_freq = n_freq;
  h->n_filter = p->n_filter;
  h->n_coef = p->n_coef;
  int64_t off = fe_align16(sizeof(fe_blob_header));
  h->off_window = (int32_t)off;    off = fe_align16(off + (int64_t)n_fft * 4);
  h->off_twiddle = (int32_t)off;   off = fe_align16(off + (int64_t)nh * 8);
  h->off_rtwiddle = (int32_t)off;  off = fe_align16(off + (int64_t)(nh / 2 + 1) * 8);
  h->off_band_start = (int32_t)off; off = fe_align16(off + (int64_t)p->n_filter * 4);
  h->off_band_len = (int32_t)off;   off = fe_align16(off + (int64_t)p->n_filter * 4);
  h->off_band_woff = (int32_t)off;  off = fe_align16(off + (int64_t)p->n_filter * 4);
  // Band extents: with a filterbank given, the exact extents; without (size query), the worst case.
  int64_t total_w = 0;
  int max_len = 0;
  if (fbank) {
    starts->assign(p->n_filter, 0);
    lens->assign(p->n_filter, 0);
    for (int f = 0; f < p->n_filter; ++f) {
      int first = -1, last = -1;
      for (int k = 0; k < n_freq; ++k) {
        if (fbank[(int64_t)k * p->n_filter + f] != 0.0f) {
          if (first < 0) first = k;
          last = k;
        }
      }
      if (first >= 0) {
        (*starts)[f] = first;
        (*lens)[f] = last - first + 1;
      }
      total_w += (*lens)[f];
      if ((*lens)[f] > max_len) max_len = (*lens)[f];
    }
  } else {
    total_w = (int64_t)p->n_filter * n_freq;
    max_len = n_freq;
  }
  h->off_band_w = (int32_t)off;
  h->total_w = (int32_t)total_w;
  h->max_band_len = max_len;
  // The size query must not depend on the filterbank contents: always reserve the dense worst case.
  off = fe_align16(off + (int64_t)p->n_filter * n_freq * 4);
  h->off_dct = (int32_t)off;
  off = fe_align16(off + (int64_t)p->n_filter * (p->n_coef > 0 ? p->n_coef : 0) * 4);
  off = fe_gemm_plan_layout(p, h, off);
  if (off > 0x7fffffff) { fe_set_error("tables blob too large"); return B200FE_ERR_UNSUPPORTED; }
  h->total_bytes = (int32_t)off;
  return B200FE_OK;
}

extern "C" int64_t b200fe_tables_bytes(const b200fe_params* p) {
  int32_t st = fe_validate_params(p);
  if (st != B200FE_OK) return st;
  fe_blob_header h;
  st = plan_layout(p, nullptr, &h, nullptr, nullptr);
  if (st != B200FE_OK) return st;
  return h.total_bytes;
}

extern "C" int32_t b200fe_tables_pack(const b200fe_params* p, const float* window, const float* fbank,
                                      const float* dct, void* blob_host, size_t blob_bytes) {
  int32_t st = fe_validate_params(p);
  if (st != B200FE_OK) return st;
  if (!window || !blob_host) { fe_set_error("window / blob_host is NULL"); return B200FE_ERR_BAD_ARG; }
  if (p->n_filter > 0 && !fbank) { fe_set_error("fbank is NULL but n_filter=%d", p->n_filter); return B200FE_ERR_BAD_ARG; }
  if (p->n_coef > 0 && (!dct || p->n_filter == 0)) { fe_set_error("dct is NULL (or n_filter=0) but n_coef=%d", p->n_coef); return B200FE_ERR_BAD_ARG; }
  fe_blob_header h;
  std::vector<int32_t> starts, lens;
  st = plan_layout(p, p->n_filter > 0 ? fbank : nullptr, &h, &starts, &lens);
  if (st != B200FE_OK) return st;
  if (p->n_filter == 0) { h.total_w = 0; h.max_band_len = 0; }
  if (blob_bytes < (size_t)h.total_bytes) {
    fe_set_error("tables blob: %zu bytes given, %d needed", blob_bytes, h.total_bytes);
    return B200FE_ERR_WORKSPACE;
  }
  char* base = (char*)blob_host;
  memset(base, 0, h.total_bytes);
  const int n_fft = p->n_fft, nh = n_fft / 2, n_freq = nh + 1;
  // window, centred like torch.stft (left pad (n_fft - win_length) / 2)
  float* w = (float*)(base + h.off_window);
  const int left = (n_fft - p->win_length) / 2;
  for (int i = 0; i < p->win_length; ++i) w[left + i] = window[i];
  // twiddles in double, rounded once
  fe_c2* tw = (fe_c2*)(base + h.off_twiddle);
  for (int k = 0; k < nh; ++k) {
    double a = -2.0 * M_PI * (double)k / (double)nh;
    tw[k].x = (float)cos(a);
    tw[k].y = (float)sin(a);
  }
  fe_c2* rtw = (fe_c2*)(base + h.off_rtwiddle);
  for (int k = 0; k <= nh / 2; ++k) {
    double a = -2.0 * M_PI * (double)k / (double)n_fft;
    rtw[k].x = (float)cos(a);
    rtw[k].y = (float)sin(a);
  }
  if (p->n_filter > 0) {
    int32_t* bs = (int32_t*)(base + h.off_band_start);
    int32_t* bl = (int32_t*)(base + h.off_band_len);
    int32_t* bo = (int32_t*)(base + h.off_band_woff);
    float* bw = (float*)(base + h.off_band_w);
    int32_t woff = 0;
    for (int f = 0; f < p->n_filter; ++f) {
      bs[f] = starts[f];
      bl[f] = lens[f];
      bo[f] = woff;
      for (int i = 0; i < lens[f]; ++i) bw[woff + i] = fbank[(int64_t)(starts[f] + i) * p->n_filter + f];
      woff += lens[f];
    }
  }
  if (p->n_coef > 0) memcpy(base + h.off_dct, dct, (size_t)p->n_filter * p->n_coef * 4);
  st = fe_gemm_pack(p, &h, window, fbank, base);
  if (st != B200FE_OK) return st;
  memcpy(base, &h, sizeof(h));
  (void)n_freq;
  return B200FE_OK;
}

extern "C" int32_t b200fe_tables_variant(const b200fe_params* p, const void* blob_host) {
  int32_t st = fe_validate_params(p);
  if (st != B200FE_OK) return st;
  if (!blob_host) { fe_set_error("blob_host is NULL"); return B200FE_ERR_BAD_ARG; }
  const fe_blob_header* h = (const fe_blob_header*)blob_host;
  if (h->magic != FE_BLOB_MAGIC) { fe_set_error("blob_host is not a packed tables blob"); return B200FE_ERR_BAD_ARG; }
  const bool gemm_ok = h->gemm_ok != 0 && fe_gemm_variant_built();
  if (p->variant == B200FE_VARIANT_FFT) return B200FE_VARIANT_FFT;
  if (p->variant == B200FE_VARIANT_DFT_GEMM) {
    if (!gemm_ok) {
      fe_set_error("variant dft_gemm is not available for this configuration (needs win_length == 2*hop_length, "
                   "n_fft <= 512, symmetric window with window[0] == 0, triangular filterbank, n_filter <= 32)");
      return B200FE_ERR_UNSUPPORTED;
    }
    return B200FE_VARIANT_DFT_GEMM;
  }
  return (gemm_ok && fe_gemm_auto_prefers(p)) ? B200FE_VARIANT_DFT_GEMM : B200FE_VARIANT_FFT;
}
