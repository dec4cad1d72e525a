// Per-thread arithmetic of the DFT-GEMM variant (see fe_gemm_layout.h for the math and layouts).
// Everything here compiles for the device (used by fe_stream.cu) and as plain C++ (tests/emu), so the
// fold / scale / fp16-split / filterbank-sweep logic is exercised on the CPU with the MMA replaced by
// loops over the very same operand images.
#ifndef FE_GEMM_CUH_
#define FE_GEMM_CUH_

#include <math.h>
#include <string.h>
#include <cuda_fp16.h>

#include "fe_gemm_layout.h"

struct alignas(16) fe_u4 {
  uint32_t x, y, z, w;
};

FE_HD uint32_t fe_f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
FE_HD float fe_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

// Two floats -> packed fp16 pair (first in the low half) with round-to-nearest, and the residuals
// v - fp16(v) (exact in fp32) for the lo parts.
FE_HD uint32_t fe_pack_hi(float a, float b, float& ra, float& rb) {
#ifdef __CUDA_ARCH__
  const __half2 h = __floats2half2_rn(a, b);
  const float2 back = __half22float2(h);
  ra = a - back.x;
  rb = b - back.y;
  return *reinterpret_cast<const uint32_t*>(&h);
#else
  const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
  ra = a - __half2float(ha);
  rb = b - __half2float(hb);
  uint16_t ua, ub;
  memcpy(&ua, &ha, 2);
  memcpy(&ub, &hb, 2);
  return (uint32_t)ua | ((uint32_t)ub << 16);
#endif
}
FE_HD uint32_t fe_pack_lo(float a, float b) {
  float ra, rb;
  return fe_pack_hi(a, b, ra, rb);
}

// Per-frame power-of-two scale: `bound` >= max |a_e|, |a_o| of the frame (2 * max|x| over its two hop
// blocks).  scale * bound lies in [2^13, 2^14), so the fp16 hi parts keep 11 significant bits and the lo
// parts stay normal numbers; `unscale` = 1 / (scale * 2^14) undoes it (and the 2^14 of the DFT tiles)
// on the amplitude, i.e. power_true = power_acc * unscale^2.
FE_HD void fe_gemm_frame_scale(float bound, float& scale, float& unscale) {
  const uint32_t eb = (fe_f2u(bound) >> 23) & 0xffu;
  int e = (int)eb - 127;               // bound in [2^e, 2^(e+1))
  if (eb == 0u || eb == 0xffu) e = FE_GEMM_A_SCALE_LOG2;   // zero / subnormal / non-finite: scale 1
  int s = FE_GEMM_A_SCALE_LOG2 - e;
  s = s > 100 ? 100 : (s < -100 ? -100 : s);
  scale = fe_u2f((uint32_t)(s + 127) << 23);
  unscale = fe_u2f((uint32_t)(-s - FE_GEMM_B_SCALE_LOG2 + 127) << 23);
}

// =====================================================================================================
// Streaming kernel (fe_stream.cu)
// =====================================================================================================

// Tile geometry of the frame stream.  The launch's rows are one stream of frames g = row * nF + t; a tile takes
// `tile_frames` consecutive stream frames.  Hop block v of a row (samples [(v-1) hop, v hop); v = 0 and v = nF are
// the reflect-padded edges) has stream index row (nF+1) + v; frame (row, t) reads blocks t (backward half) and
// t + 1 (forward half), so a tile needs the nv consecutive stream blocks sv0 .. sv0 + nv - 1.
struct fe_tile_geo {
  int g0, count, row0, row_last, sv0, nv;
};
FE_HD fe_tile_geo fe_tile_geometry(int tile, int tile_frames, int total_frames, int nF) {
  fe_tile_geo t;
  t.g0 = tile * tile_frames;
  t.count = total_frames - t.g0 < tile_frames ? total_frames - t.g0 : tile_frames;
  t.row0 = t.g0 / nF;
  const int g_last = t.g0 + t.count - 1;
  t.row_last = g_last / nF;
  t.sv0 = t.g0 + t.row0;
  t.nv = t.count + (t.row_last - t.row0) + 1;
  return t;
}
// frames per tile: 128 (the TMEM lanes) unless the utterances are so short that 128 frames would span more than
// three of them (the sample buffer holds 132 hop blocks)
FE_HD int fe_tile_frames(int nF) { return 2 * nF < FE_GEMM_TILE_M ? 2 * nF : FE_GEMM_TILE_M; }

// One production unit = 16 sample pairs j = j0 .. j0+15 of one frame.  The
// scale is folded into the fold:  bs = b*s ; a_e*s = fma(f, s, bs) ; a_o*s = fma(f, s, -bs)  (s is a power of
// two, so both are exactly (f +- b)*s) and bin n_fft/4 is accumulated from the SCALED values with the
// interleaved weight table midc[j] = (j even ? Re weight : Im weight) (four 16-byte broadcast loads per unit).
FE_HD void fe_stream_produce_unit(const float* fwd, const float* bwd, float scale, const float* midc,
                                  float& mid_re, float& mid_im, fe_u4* chunk) {
  uint32_t hi[4][4], lo[4][4];
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int i0 = 4 * w;
    float ae[4], ao[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float bs = bwd[i0 + u] * scale;
      ae[u] = fmaf(fwd[i0 + u], scale, bs);
      ao[u] = fmaf(fwd[i0 + u], scale, -bs);
    }
    const float m0 = midc[i0], m1 = midc[i0 + 1], m2 = midc[i0 + 2], m3 = midc[i0 + 3];
    mid_re = fmaf(ae[0], m0, mid_re);
    mid_im = fmaf(ao[1], m1, mid_im);
    mid_re = fmaf(ae[2], m2, mid_re);
    mid_im = fmaf(ao[3], m3, mid_im);
    float r0, r1;
    hi[0][w] = fe_pack_hi(ae[0], ae[2], r0, r1); lo[0][w] = fe_pack_lo(r0, r1);
    hi[1][w] = fe_pack_hi(ae[1], ae[3], r0, r1); lo[1][w] = fe_pack_lo(r0, r1);
    hi[2][w] = fe_pack_hi(ao[0], ao[2], r0, r1); lo[2][w] = fe_pack_lo(r0, r1);
    hi[3][w] = fe_pack_hi(ao[1], ao[3], r0, r1); lo[3][w] = fe_pack_lo(r0, r1);
  }
#pragma unroll
  for (int sub = 0; sub < 4; ++sub) {
    chunk[sub * 2 + 0] = fe_u4{hi[sub][0], hi[sub][1], hi[sub][2], hi[sub][3]};
    chunk[sub * 2 + 1] = fe_u4{lo[sub][0], lo[sub][1], lo[sub][2], lo[sub][3]};
  }
}

// ---- drain: sliding even/odd filter accumulators (tables: fe_gemm_layout.h) ------------------------------
struct fe_drain_state {
  float acc[4];
  int off[4];   // element offset of the accumulator's filter row in the emission scratch (dummy row: none)
};

// adds one finished accumulator to the frame's filter sum (e_col = this frame's column of the [filter + 1][128] array;
// accumulators without a filter point at the dummy last row, so there is nothing to test)
FE_HD void fe_drain_emit(float* e_col, int off, float v, float us2) { e_col[off] = fmaf(v, us2, e_col[off]); }

// the (rare, thread-uniform) switches of one column
FE_HD void fe_drain_switch(unsigned flags, fe_drain_ids ids, fe_drain_state& st, float* e_col, float us2) {
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (flags & (1u << a)) {
      fe_drain_emit(e_col, st.off[a], st.acc[a], us2);
      st.acc[a] = 0.0f;
      st.off[a] = ids.off[a];
    }
  }
}

// NB consecutive columns starting at a multiple of 8 (ctl = the batch's switch word, w / ids at the first column)
template <int NB>
FE_HD void fe_drain_cols(const fe_drain_w* w, const fe_drain_ids* ids, unsigned ctl, const float* ce, const float* co,
                         const float* se, const float* so, fe_drain_state& st, float* e_col, float us2) {
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const float re1 = ce[i] + co[i], im1 = se[i] + so[i], re2 = ce[i] - co[i], im2 = so[i] - se[i];
    const float p1 = fmaf(re1, re1, im1 * im1);  // |X[k]|^2 (scaled units)
    const float p2 = fmaf(re2, re2, im2 * im2);  // |X[n_fft/2 - k]|^2
    const unsigned fl = (ctl >> (4 * i)) & 15u;
    if (fl) fe_drain_switch(fl, ids[i], st, e_col, us2);
    const fe_drain_w t = w[i];
    st.acc[0] = fmaf(p1, t.w[0], st.acc[0]);
    st.acc[1] = fmaf(p1, t.w[1], st.acc[1]);
    st.acc[2] = fmaf(p2, t.w[2], st.acc[2]);
    st.acc[3] = fmaf(p2, t.w[3], st.acc[3]);
  }
}

// bin n_fft/4 (column index nhalf of the tables): only the lo-run accumulators
FE_HD void fe_drain_mid(const fe_drain_w* w, const fe_drain_ids* ids, unsigned ctl, float p_mid, fe_drain_state& st,
                        float* e_col, float us2) {
  const unsigned fl = ctl & 3u;
  if (fl) fe_drain_switch(fl, ids[0], st, e_col, us2);
  st.acc[0] = fmaf(p_mid, w[0].w[0], st.acc[0]);
  st.acc[1] = fmaf(p_mid, w[0].w[1], st.acc[1]);
}

FE_HD void fe_drain_flush(fe_drain_state& st, float* e_col, float us2) {
#pragma unroll
  for (int a = 0; a < 4; ++a) fe_drain_emit(e_col, st.off[a], st.acc[a], us2);
}

#endif  // FE_GEMM_CUH_
