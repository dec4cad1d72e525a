// Shared definitions of the B200 spectral front-end library (host + device).
#ifndef FE_COMMON_H_
#define FE_COMMON_H_

#include <stdint.h>
#include <stddef.h>
#include "b200fe.h"

#if defined(__CUDACC__)
#define FE_HD __host__ __device__ __forceinline__
#else
#define FE_HD inline
#endif

// 8-byte complex; LDS.64 / STS.64 on the device, plain struct in the CPU emulation build.
struct alignas(8) fe_c2 {
  float x, y;
};

#define FE_BLOB_MAGIC 0xB200FE03u

// Header of the constant-table blob produced by b200fe_tables_pack (all offsets in bytes from the
// start of the blob, 16-byte aligned).  The blob is position independent: the same bytes are valid
// on the host and on the device.
struct fe_blob_header {
  uint32_t magic;
  int32_t abi_version;
  int32_t n_fft, win_length, hop_length, n_freq, n_filter, n_coef;
  int32_t total_bytes;
  int32_t off_window;      // float[n_fft]        window zero-padded centred to n_fft (torch.stft)
  int32_t off_twiddle;     // fe_c2[n_fft/2]      exp(-2*pi*i*k/(n_fft/2))
  int32_t off_rtwiddle;    // fe_c2[n_fft/4+1]    exp(-2*pi*i*k/n_fft)
  int32_t off_band_start;  // int32[n_filter]     first non-zero bin of each filter
  int32_t off_band_len;    // int32[n_filter]     number of bins from first to last non-zero
  int32_t off_band_woff;   // int32[n_filter]     offset (floats) of the filter's weights
  int32_t off_band_w;      // float[total_w]
  int32_t total_w;
  int32_t max_band_len;
  int32_t off_dct;         // float[n_filter][n_coef]
  // ---- DFT-GEMM variant (0 when the configuration does not support it) ----
  int32_t gemm_ok;         // 1 when the tiles below are present
  int32_t gemm_kpairs;     // folded K rows per parity class (multiple of 16)
  int32_t gemm_nhalf;      // n_fft/4 : GEMM N (bins 0 .. n_fft/4-1; bin n_fft/4 handled apart)
  int32_t off_gemm_b;      // __half operand tiles, see fe_gemm_layout.h
  int32_t gemm_b_bytes;
  int32_t off_gemm_mid;    // float[2][gemm_kpairs]  true-unit weights of bin n_fft/4 (Re from a_e, Im from a_o)
  // sliding even/odd filter accumulators of the drain (fe_gemm_layout.h)
  int32_t off_gemm_dw;     // fe_drain_w[2 runs][nhalf/2]        per column pair: weights of the segment active at the batch start
  int32_t off_gemm_dctl;   // uint32[2 runs][nhalf/8]            per batch: boundary flags and new targets
  int32_t off_gemm_dids;   // fe_drain_hdr                       first / last targets of the runs, straddler merge flags
  int32_t off_gemm_dwn;    // float[2 runs][nhalf/8][2][4][2]    per batch and class: weights behind the boundary (pair, half)
  int32_t reserved[6];
};

static inline int64_t fe_align16(int64_t v) { return (v + 15) & ~(int64_t)15; }

// Validates params; returns B200FE_OK or an error, writing a message through fe_set_error.
int32_t fe_validate_params(const b200fe_params* p);
void fe_set_error(const char* fmt, ...);

#endif  // FE_COMMON_H_
