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

// Packed fp32x2 arithmetic (sm_100 FADD2 / FMUL2 / FFMA2: one issue slot for two lanes); plain C++ on the host.
struct fe_f2 {
  float x, y;
};
FE_HD fe_f2 fe_add2(fe_f2 a, fe_f2 b) {
#ifdef __CUDA_ARCH__
  const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
  return fe_f2{r.x, r.y};
#else
  return fe_f2{a.x + b.x, a.y + b.y};
#endif
}
FE_HD fe_f2 fe_mul2(fe_f2 a, fe_f2 b) {
#ifdef __CUDA_ARCH__
  const float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
  return fe_f2{r.x, r.y};
#else
  return fe_f2{a.x * b.x, a.y * b.y};
#endif
}
FE_HD fe_f2 fe_fma2(fe_f2 a, fe_f2 b, fe_f2 c) {
#ifdef __CUDA_ARCH__
  const float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
  return fe_f2{r.x, r.y};
#else
  return fe_f2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)};
#endif
}

// Two floats -> packed fp16 pair (first in the low half, round to nearest) and the packed fp16 pair of the residuals
// v - fp16(v) (exact in fp32): the hi / lo operands of the split-fp16 products.
FE_HD void fe_split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
  // cvt.rn.f16x2 for the hi pair, then the residuals with the sm_100 mixed-precision FMA (FHFMA: fp16 x fp16 + fp32 ->
  // fp32): fma(h, -1, v) = v - h, exact in fp32 -- one instruction per value instead of unpack + subtract
  uint32_t h, l;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
  float ra, rb;
  asm("{\n\t.reg .b16 h0, h1, m1;\n\tmov.b32 {h0, h1}, %2;\n\tmov.b16 m1, 0xBC00;\n\t"
      "fma.rn.f32.f16 %0, h0, m1, %3;\n\tfma.rn.f32.f16 %1, h1, m1, %4;\n\t}\n"
      : "=f"(ra), "=f"(rb) : "r"(h), "f"(a), "f"(b));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(rb), "f"(ra));
  hi = h;
  lo = l;
#else
  float ra, rb;
  hi = fe_pack_hi(a, b, ra, rb);
  lo = fe_pack_lo(ra, rb);
#endif
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
#ifdef __CUDA_ARCH__
    // the same fused multiply-adds, two lanes per issue slot (FMUL2 / FFMA2)
#pragma unroll
    for (int u = 0; u < 4; u += 2) {
      const fe_f2 s2 = fe_f2{scale, scale};
      const fe_f2 f2 = fe_f2{fwd[i0 + u], fwd[i0 + u + 1]};
      const fe_f2 bs = fe_mul2(fe_f2{bwd[i0 + u], bwd[i0 + u + 1]}, s2);
      const fe_f2 e = fe_fma2(f2, s2, bs);
      const fe_f2 o = fe_fma2(f2, s2, fe_f2{-bs.x, -bs.y});
      ae[u] = e.x; ae[u + 1] = e.y;
      ao[u] = o.x; ao[u + 1] = o.y;
    }
#else
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float bs = bwd[i0 + u] * scale;
      ae[u] = fmaf(fwd[i0 + u], scale, bs);
      ao[u] = fmaf(fwd[i0 + u], scale, -bs);
    }
#endif
    const float m0 = midc[i0], m1 = midc[i0 + 1], m2 = midc[i0 + 2], m3 = midc[i0 + 3];
    mid_re = fmaf(ae[0], m0, mid_re);
    mid_im = fmaf(ao[1], m1, mid_im);
    mid_re = fmaf(ae[2], m2, mid_re);
    mid_im = fmaf(ao[3], m3, mid_im);
    fe_split_pair(ae[0], ae[2], hi[0][w], lo[0][w]);
    fe_split_pair(ae[1], ae[3], hi[1][w], lo[1][w]);
    fe_split_pair(ao[0], ao[2], hi[2][w], lo[2][w]);
    fe_split_pair(ao[1], ao[3], hi[3][w], lo[3][w]);
  }
#pragma unroll
  for (int sub = 0; sub < 4; ++sub) {
    chunk[sub * 2 + 0] = fe_u4{hi[sub][0], hi[sub][1], hi[sub][2], hi[sub][3]};
    chunk[sub * 2 + 1] = fe_u4{lo[sub][0], lo[sub][1], lo[sub][2], lo[sub][3]};
  }
}

// ---- drain: sliding even/odd filter accumulators over column pairs (tables and protocol: fe_gemm_layout.h) ---------
struct fe_drain_state {
  fe_f2 acc[2];    // class = filter parity: .x even columns, .y odd columns
  int tgt[2];      // filter the class is aimed at (FE_DRAIN_NONE: none)
  float defer[2];  // half 1: right part of the segment that straddles the halves (stored once half 0's part is known)
};

FE_HD void fe_drain_init(fe_drain_state& st, const fe_drain_hdr& hdr, int run, int half) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    st.acc[c] = fe_f2{0.0f, 0.0f};
    st.tgt[c] = hdr.first[run][half][c];
    st.defer[c] = 0.0f;
  }
}

// Powers of one column pair of run RUN: |X[k]|^2 = (ce+co)^2 + (se+so)^2 (run 0) or
// |X[n_fft/2 - k]|^2 = (ce-co)^2 + (so-se)^2 (run 1), two columns at a time.
template <int RUN>
FE_HD fe_f2 fe_drain_power(fe_f2 c0, fe_f2 c1, fe_f2 s0, fe_f2 s1) {
  const fe_f2 neg = fe_f2{-1.0f, -1.0f};
  const fe_f2 re = RUN == 0 ? fe_add2(c0, c1) : fe_fma2(c1, neg, c0);
  const fe_f2 im = RUN == 0 ? fe_add2(s0, s1) : fe_fma2(s0, neg, s1);
  return fe_fma2(re, re, fe_mul2(im, im));
}

// One batch of 4 column pairs: the classes' weighted sums, then (thread-uniform) the segment ends of the batch:
// `emit(filter, value)` receives a finished segment (scaled units), the class carries on with the sum of the columns
// behind the boundary.  wn = this batch's [class][pair][half] weights behind the boundaries.
template <class Emit>
FE_HD void fe_drain_batch(const fe_f2* pw, const fe_drain_w* w, const float* wn, unsigned ctl, fe_drain_state& st, Emit& emit) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    st.acc[0] = fe_fma2(pw[p], fe_f2{w[p].w[0][0], w[p].w[0][1]}, st.acc[0]);
    st.acc[1] = fe_fma2(pw[p], fe_f2{w[p].w[1][0], w[p].w[1][1]}, st.acc[1]);
  }
  if (ctl & 3u) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (ctl & (1u << c)) {
        fe_f2 nx = fe_mul2(pw[0], fe_f2{wn[c * 8 + 0], wn[c * 8 + 1]});
#pragma unroll
        for (int p = 1; p < 4; ++p) nx = fe_fma2(pw[p], fe_f2{wn[c * 8 + 2 * p], wn[c * 8 + 2 * p + 1]}, nx);
        if (ctl & (4u << c)) st.defer[c] = st.acc[c].x + st.acc[c].y;
        else emit(st.tgt[c], st.acc[c].x + st.acc[c].y);
        st.acc[c] = nx;
        st.tgt[c] = (int)((ctl >> (8 + 8 * c)) & 255u);
      }
    }
  }
}

// what a half holds for class c after its last batch (run 0, half 1: plus bin n_fft/4 with power p_mid)
FE_HD float fe_drain_leftover(const fe_drain_state& st, int c, float p_mid, float wmid) {
  return fmaf(p_mid, wmid, st.acc[c].x + st.acc[c].y);
}

// Half 1 of a run, after its walk, with half 0's leftovers l0[class]: stores the segment that straddles the halves
// (unless it is also the run's last one) and returns the run's leftovers in left[class].
template <class Emit>
FE_HD void fe_drain_join_halves(const fe_drain_state& st, const fe_drain_hdr& hdr, int run, const float* l0, float p_mid,
                                float* left, Emit& emit) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    left[c] = fe_drain_leftover(st, c, p_mid, run == 0 ? hdr.wmid[c] : 0.0f);
    if (hdr.open_tgt[run][c] != FE_DRAIN_NONE) {
      if (hdr.open_last[run][c]) left[c] += l0[c];
      else emit(hdr.open_tgt[run][c], st.defer[c] + l0[c]);
    }
  }
}

#endif  // FE_GEMM_CUH_
