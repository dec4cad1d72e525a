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
// A drain thread owns one frame (TMEM lane) and one RUN of bins: GEMM column k carries bin k (run 0, ascending) and
// bin n_fft/2 - k (run 1, descending); run 0 also takes bin n_fft/4 (evaluated by the producers) after its last
// column.  Both runs walk ALL columns 0 .. n_fft/4 - 1 in BATCHES of 8 columns = 4 pairs (2p, 2p+1), with packed
// fp32x2 arithmetic (one issue slot for two columns).  Every bin may carry at most one even-indexed and one
// odd-indexed filter (true for triangular banks), so a run needs one accumulator per filter parity (CLASS), with an
// even-column half (.x) and an odd-column half (.y).  Along a run the filter of a class is piecewise constant; a
// SEGMENT is a maximal range of columns with the same filter (or none; the host packer absorbs short filter-less
// ranges with zero weights).  A filter occupies exactly one segment of one class of a run (or the last segment of
// both runs when it straddles bin n_fft/4), so the sum a thread holds when a segment ends is the filter's FINAL
// energy for the frame: it is stored straight to the workspace (no shared-memory scratch, no second pass).
// Segment ends are handled per batch, outside the arithmetic: the packer requires at most one segment boundary per
// class inside a batch window (k0, k0+8].  Weights `w` of a batch are those of the segment active at its first column
// (zero behind the boundary); a batch with a boundary has its flag set and a second weight set `wn` for the columns
// behind the boundary: the thread sums those into a fresh accumulator, stores the finished one and carries on with the
// fresh one, aimed at the new filter.  Flags are the same for all threads: table driven, branch-uniform.
// After the last batch the two runs' leftovers are the straddling filters: run 1 hands its pair to run 0 through
// shared memory where the targets coincide.
struct alignas(16) fe_drain_w {
  float w[2][2];     // [class = filter parity][half]: weights at columns (2p, 2p+1)
};
// COLUMN HALVES.  The walk of a run is split between two threads: half 0 takes columns [0, n_fft/8), half 1 the
// rest (and, for run 0, bin n_fft/4).  Per class at most one segment straddles column n_fft/8 ("open" segment): half 0
// ends with its left part as leftover; half 1's first boundary of that class (control bit "defer") keeps the right
// part in a register instead of storing it, and adds half 0's leftover (shared memory, one named barrier per tile)
// before it stores; when half 1 has no boundary of the class at all, the open segment is also its last one and half
// 0's leftover joins half 1's.
// control word of a batch: bit 0 / 1: class 0 / 1 has a boundary in (k0, k0+8]; bit 2 / 3: that boundary's finished
// segment is the open one (deferred); bits 8-15 / 16-23: filter the class is aimed at behind the boundary
// (FE_DRAIN_NONE: no filter)
#define FE_DRAIN_NONE 255
#define FE_DRAIN_BATCH 8
struct fe_drain_hdr {
  int32_t first[2][2][2];  // [run][half][class] filter of the half's first segment (FE_DRAIN_NONE: none)
  int32_t last[2][2];      // [run][class] filter of the run's last segment
  int32_t merge[2];        // [class] 1: the last segments of the two runs are the same filter (run 0 stores the sum)
  int32_t open_tgt[2][2];  // [run][class] filter of the segment that straddles the halves (FE_DRAIN_NONE: no such segment)
  int32_t open_last[2][2]; // [run][class] 1: half 1 has no boundary of the class: the open segment is the run's last one
  float wmid[2];           // [class] weight of bin n_fft/4 (run 0, last segment)
  int32_t pad[2];
};
// table sizes per run: w[nhalf/2] pairs, wn[nhalf/8][class 2][pair-in-batch 4] (fe_drain_w.w[class] slices), ctl[nhalf/8]

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
