// xt_prep.cuh -- fp32 features [n][321][180] (feature axis contiguous) -> the fp16 time-major image the Toeplitz GEMMs of
// conv1_tc.cu / cae_enc1_tc.cu read:  xT[(column * 41 + tb)] = 8 samples x[8tb-1 .. 8tb+6][f]  (one 16-byte row).
// The output is the transpose of the input's storage order, so a direct mapping has either strided 4-byte reads or
// scattered 16-byte writes (656 B apart).  Here a block stages 32 feature columns of one utterance through shared memory:
// 128-byte coalesced row reads (optionally normalised), conflict-free 2-byte transposed stores (row pitch 676 B = 169
// words, odd), then every column leaves as one contiguous 656-byte run.
#pragma once
#include "common.cuh"
#include "layout.cuh"

namespace dfs {
namespace {

constexpr int kXpPitch = 338;   // halfs per staged column: s = t + 1 in [0, 328), +10 pad (odd word pitch)

// SPLIT: also writes xt_lo, the fp16 rounding residual of every sample (x = hi + lo up to 2^-22 |x|; "split" precision)
// NORM: FeatureNormalizer.transform (mean / sd per feature) before the conversion.
// The first version spent ~50 instructions per sample (64-bit index products, two bounds predicates and a divergence region per sample;
// ncu: 76 % of the issue slots, DRAM at 33 %).  Here a thread walks its 40 rows (t = ty, ty + 8, ... ty + 312) with one pointer
// increment per row, the shared-memory offsets are immediates, and the only predicates left are the block-uniform column bound of the
// last column block and row 320 (row group 0 only).  Same samples, same conversions, same bits.
static_assert(kT == 321, "row schedule below: 5 rounds of 8 rows per row group + row 320");
template <bool SPLIT = false, bool NORM = false>
__global__ void __launch_bounds__(256) xt_prep_transpose_kernel(const float* __restrict__ x, long long sn, long long st, int cols, int col_pad,
                                                                 int lead_rows, const float* __restrict__ mean, const float* __restrict__ sd,
                                                                 uint16_t* __restrict__ xt, uint16_t* __restrict__ xt_lo = nullptr) {
  __shared__ __align__(16) uint16_t tile[(SPLIT ? 2 : 1) * 32 * kXpPitch];
  constexpr int kLo = 32 * kXpPitch;   // offset of the residual tile
  const int n = blockIdx.y, f0 = 32 * blockIdx.x;
  const int fl = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int f = f0 + fl;
  // zero the samples outside [0, 321): s = 0 (t = -1) and s = 322 .. 327
  if (threadIdx.x < 32) {
    tile[threadIdx.x * kXpPitch] = 0;
    if constexpr (SPLIT) tile[kLo + threadIdx.x * kXpPitch] = 0;
#pragma unroll
    for (int s = kT + 1; s < 328; ++s) {
      tile[threadIdx.x * kXpPitch + s] = 0;
      if constexpr (SPLIT) tile[kLo + threadIdx.x * kXpPitch + s] = 0;
    }
  }
  const bool colok = f < kF;   // false only in the last column block (columns 180 .. 191 are padding)
  float m = 0.0f, sg = 1.0f;
  if (NORM && colok) { m = mean[f]; sg = sd[f]; }
  // one sample: normalise (FeatureNormalizer.transform, before the zero padding), one saturating convert -- |a| > 65504 and +-inf ->
  // +-65504, NaN stays NaN (fmaxf / fminf clamps would swallow it) -- and the transposed 2-byte store
  auto put = [&](float a, uint16_t* dst) {
    if constexpr (NORM) a = (a - m) / sg;   // padding columns: (0 - 0) / 1 = 0
    const uint32_t hi = pack_act2(a, 0.0f);
    dst[0] = (uint16_t)(hi & 0xffffu);
    if constexpr (SPLIT) dst[kLo] = (uint16_t)(pack_act2_residual(a, 0.0f, hi) & 0xffffu);
  };
  const float* p = x + (long long)n * sn + (long long)ty * st + f;   // dereferenced only where colok
  const long long step = 8 * st;
  uint16_t* trow = tile + fl * kXpPitch + ty + 1;
#pragma unroll 1
  for (int r = 0; r < 5; ++r) {   // 8 independent row loads in flight per thread
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = colok ? __ldg(p) : 0.0f;
      p += step;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) put(v[k], trow + 8 * k);
    trow += 64;
  }
  if (ty == 0) put(colok ? __ldg(p) : 0.0f, trow);   // row 320: p and trow have advanced 40 rows
  __syncthreads();
  const int nf = (kF - f0) < 32 ? (kF - f0) : 32;
  for (int item = threadIdx.x; item < nf * 41; item += 256) {
    const int c = item / 41, tb = item - c * 41;
    const uint32_t* p4 = reinterpret_cast<const uint32_t*>(tile + c * kXpPitch + 8 * tb);
    uint16_t* dst = xt + ((long long)lead_rows + ((long long)n * cols + f0 + c + col_pad) * 41 + tb) * 8;
    st_global_v4(dst, p4[0], p4[1], p4[2], p4[3]);
    if constexpr (SPLIT) {
      const uint32_t* pl = reinterpret_cast<const uint32_t*>(tile + kLo + c * kXpPitch + 8 * tb);
      st_global_v4(xt_lo + (dst - xt), pl[0], pl[1], pl[2], pl[3]);
    }
  }
}

}  // namespace
}  // namespace dfs
