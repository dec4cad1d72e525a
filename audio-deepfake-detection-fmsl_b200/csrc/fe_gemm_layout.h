// Data layout shared by the host packer (fe_gemm_tables.cpp), the tcgen05 kernel (fe_stream.cu) and
// the CPU emulation (tests/emu) of the DFT-GEMM variant.
//
// Math.  A frame is win = 2*hop samples centred on c (torch.stft center=True); with the symmetric
// window w(j) = w(-j) at offset j from the centre and the folded samples
//     a_e[j] = x[c+j] + x[c-j],   a_o[j] = x[c+j] - x[c-j],   j = 0 .. win/2-1   (a_e[0] = 2 x[c])
// the spectrum is  Re X[k] = sum_j a_e[j] h(j) w(j) cos(2 pi k j / n_fft),  h(0) = 1/2, h(j>0) = 1
//                  Im X[k] = - sum_j a_o[j] w(j) sin(2 pi k j / n_fft)      (up to the sign (-1)^k).
// Splitting j by parity gives four sums  ce, co (cos, even / odd j)  and  se, so (sin, even / odd j)
// with  X[k] = (ce+co) + i(se+so)  and  X[n_fft/2 - k] = (ce-co) + i(so-se), so bins 0..n_fft/4-1
// (columns of the GEMM) yield all bins except n_fft/4, which the producer evaluates directly.
// Each of the four sums is a GEMM  [128 frames x 16*stages] x [16*stages x n_fft/4]  evaluated as three
// fp16 products  A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  accumulated in fp32 in TMEM.
#ifndef FE_GEMM_LAYOUT_H_
#define FE_GEMM_LAYOUT_H_

#include "fe_common.h"

#define FE_GEMM_MAX_FILTERS 32
#define FE_GEMM_TILE_M 128        // frames per tile = TMEM lanes
#define FE_GEMM_B_SCALE_LOG2 14   // DFT matrix entries are stored times 2^14 (lo parts stay normal fp16)
#define FE_GEMM_A_SCALE_LOG2 13   // per frame: 2*max|x| is scaled into [2^13, 2^14)
#define FE_GEMM_STAGE_J 32        // sample pairs per pipeline stage: one K=16 MMA step per sub-GEMM

// ---- drain tables ----------------------------------------------------------------------------------------
// A drain thread owns one frame (TMEM lane) and one of FE_DRAIN_GROUPS column groups (nhalf / 4 consecutive GEMM
// columns).  Column k carries bin k (the ascending "lo" run) and bin n_fft/2 - k (the descending "hi" run).  Every
// bin may have at most one even-indexed and one odd-indexed filter with non-zero weight (true for triangular
// banks), so each run needs two accumulators: class a = 2*run + parity.  Columns are processed in PAIRS (2p, 2p+1)
// with packed fp32x2 arithmetic (one issue slot for two columns), so every class has two independent halves h (even /
// odd column), each with its own target filter: half (a, h) is emitted (added to the frame's filter sum) and
// re-targeted before pair p whenever the filter of class a at column 2p + h differs from the one at column 2p + h - 2
// ("switch").  Switches are the same for all threads: table driven and branch-uniform.
struct alignas(16) fe_drain_w {
  float w[4][2];     // per class a: weights at columns (2p, 2p+1); classes 0,1: bin k (even, odd filter); 2,3: bin n_fft/2 - k
};
struct alignas(16) fe_drain_ids {
  int16_t off[8];    // per half 2a + h, once this pair's switches are done: BYTE offset (filter * 128 * 4) of its filter's
                     // row in the frame-major emission scratch; n_filter * 512 (a dummy row) while it has none.  The first
                     // pair of a column group carries no switch flags: the walk starts from that pair's entry.
};
#define FE_DRAIN_GROUPS 4   // column groups (warps per TMEM lane quarter)

// UMMA K-major, no-swizzle operand tile of `rows` rows x 16 K-values (one K=16 MMA step):
// [K chunk of 8][row][8 halfs] -> descriptor LBO (K-chunk stride) = rows*16 B, SBO (8-row group) = 128 B.
FE_HD int fe_gemm_operand_offset(int rows, int r, int kk) { return (kk >> 3) * rows * 16 + r * 16 + (kk & 7) * 2; }
FE_HD int fe_gemm_tile_bytes(int rows) { return 2 * rows * 16; }
// a stage holds [sub-GEMM 4: ce co se so][flavour 2: hi lo] tiles
FE_HD int fe_gemm_b_tile_offset(int nhalf, int sub, int flav) { return (sub * 2 + flav) * fe_gemm_tile_bytes(nhalf); }
FE_HD int fe_gemm_b_stage_bytes(int nhalf) { return 8 * fe_gemm_tile_bytes(nhalf); }
FE_HD int fe_gemm_a_tile_offset(int sub, int flav) { return (sub * 2 + flav) * fe_gemm_tile_bytes(FE_GEMM_TILE_M); }
FE_HD int fe_gemm_a_stage_bytes() { return 8 * fe_gemm_tile_bytes(FE_GEMM_TILE_M); }

#endif  // FE_GEMM_LAYOUT_H_
