// layout.cuh -- the activation layout shared by the CUDA-core producers and the tcgen05 conv kernels.
//
// "FT8" padded planar layout, fp16:
//     elem(plane j, column gc, row t', e) at  ((j * ncols + gc) * RS + t') * 8 + e
//   * channel c = 8*j + e            -- a "plane" holds 8 channels = one 16-byte K chunk of an MMA
//   * column  gc = n * COLS + f'     -- utterance n, padded feature index f' in [0, F+1], COLS = F+2
//   * row     t' in [0, T+1]         -- padded time index, RS = T+2
//   * f' = 0, f' = F+1, t' = 0, t' = T+1 are zero padding (the conv's padding=1); they are zeroed
//     once when the workspace is created and never written afterwards.
//
// Why this shape: with time fastest and 16-byte channel chunks, 8 consecutive time steps of one
// feature column form exactly one 8x16B UMMA "core matrix" of the SWIZZLE_NONE K-major canonical
// layout, the next feature column is a constant byte stride away (the descriptor's SBO), and a
// 3x3 tap (dt, df) is nothing but a constant byte offset of the descriptor's start address.  One
// TMA box load of (8*MT+2) rows x 18 columns x all planes therefore feeds all 9 taps of MT
// 128-row MMA tiles (16 feature columns x 8 time steps each) with no im2col and no re-load.
#pragma once
#include <stdint.h>

namespace dfs {

constexpr int kT = 321;       // input frames
constexpr int kF = 180;       // input features
constexpr int kCols = kF + 2; // padded feature columns per utterance (CNN2D keeps F through all layers)
constexpr int kColTile = 16;  // feature columns per MMA tile
constexpr int kRowTile = 8;   // time steps per MMA tile

struct ActBuf {
  uint16_t* ptr;   // fp16 bits
  int planes;      // C / 8
  int RS;          // T + 2
  int64_t ncols;   // allocated columns per plane (n_max * COLS + slack)
  __host__ __device__ int64_t plane_elems() const { return ncols * RS * 8; }
  __host__ __device__ int64_t bytes() const { return plane_elems() * planes * 2; }
};

// column tiles needed to cover gc in [1, n*COLS)
__host__ __device__ inline int num_col_tiles(int64_t n, int cols) { return (int)((n * cols - 1 + kColTile - 1) / kColTile); }

}  // namespace dfs
