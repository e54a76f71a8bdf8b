// layout.cuh -- the activation layouts shared by the producers and the tcgen05 conv kernels.
//
// "FT8" padded planar layout, fp16 (act2, the input of conv3):
//     elem(plane j, column gc, row t', e) at  ((j * ncols + gc) * RS + t') * 8 + e
//   * channel c = 8*j + e            -- a "plane" holds 8 channels = one 16-byte K chunk of an MMA
//   * column  gc = n * COLS + f'     -- utterance n, padded feature index f' in [0, F+1], COLS = F+2
//   * row     t' in [0, T+1]         -- padded time index, RS = T+2
//   * f' = 0, f' = F+1, t' = 0, t' = T+1 are zero padding (the conv's padding=1); they are zeroed
//     once when the workspace is created and never written afterwards.
//
// "FT8P" = FT8 with the time axis split by parity (act1, the input of conv2): time step t (0..159) is
// stored at s = t + 2 = 2*row + par, i.e. plane index par*4 + j, row = t/2 + 1, par = t&1; RS = 82.
// A GEMM row of conv2 is a PAIR of output time steps, so that the two conv outputs that the (2,1)
// average pool combines are two column blocks of the same accumulator row (conv_tc.cu, PAIR mode).
//
// Why this shape: with time fastest and 16-byte channel chunks, 8 consecutive rows of one feature
// column form exactly one 8x16B UMMA "core matrix" of the SWIZZLE_NONE K-major canonical layout, the
// next feature column is a constant byte stride away (the descriptor's SBO), and a 3x3 tap is
// nothing but a constant byte offset of the descriptor's start address.  One TMA box load of
// (8*MT+2) rows x 18 columns x all planes therefore feeds all taps of MT 128-row MMA tiles
// (16 feature columns x 8 rows each) with no im2col and no re-load.
#pragma once
#include <stdint.h>

namespace dfs {

constexpr int kT = 321;       // input frames
constexpr int kF = 180;       // input features
constexpr int kCols = kF + 2; // padded feature columns per utterance (CNN2D keeps F through all layers)
constexpr int kColTile = 16;  // feature columns per MMA tile
constexpr int kRowTile = 8;   // rows per MMA tile
constexpr int kAct1RS = 82;   // rows per column of act1 (FT8P: 80 time pairs + 2 pads), 8 planes
constexpr int kAct2RS = 82;   // rows per column of act2 (FT8: 80 time steps + 2 pads), 8 planes

struct ActBuf {
  uint16_t* ptr;   // fp16 bits
  int planes;      // C / 8 (x2 for FT8P)
  int RS;          // rows per column incl. padding
  int64_t ncols;   // allocated columns per plane (n_max * COLS + slack)
  __host__ __device__ int64_t plane_elems() const { return ncols * RS * 8; }
  __host__ __device__ int64_t bytes() const { return plane_elems() * planes * 2; }
};

// column tiles needed to cover gc in [1, n*COLS)
__host__ __device__ inline int num_col_tiles(int64_t n, int cols) { return (int)((n * cols - 1 + kColTile - 1) / kColTile); }

}  // namespace dfs
