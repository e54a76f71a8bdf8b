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
template <bool SPLIT = false>
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
  float m = 0.0f, sg = 1.0f;
  if (mean != nullptr && f < kF) { m = mean[f]; sg = sd[f]; }
  const float* src = x + (long long)n * sn + f;
  // 8 independent row loads in flight per thread (the plain loop issued one load, waited, converted, stored: ncu showed the
  // kernel at 15 % of the DRAM rate with every warp parked on the long scoreboard)
  for (int t0 = ty; t0 < kT; t0 += 64) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int t = t0 + 8 * k;
      v[k] = (f < kF && t < kT) ? __ldg(src + (long long)t * st) : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int t = t0 + 8 * k;
      float a = v[k];
      if (mean != nullptr) a = (a - m) / sg;     // FeatureNormalizer.transform, before the zero padding
      a = (f < kF) ? a : 0.0f;
      // one saturating convert: |a| > 65504 and +-inf -> +-65504, NaN stays NaN (fmaxf / fminf clamps would swallow it)
      if (t < kT) {
        const uint32_t hi = pack_act2(a, 0.0f);
        tile[fl * kXpPitch + t + 1] = (uint16_t)(hi & 0xffffu);
        if constexpr (SPLIT) tile[kLo + fl * kXpPitch + t + 1] = (uint16_t)(pack_act2_residual(a, 0.0f, hi) & 0xffffu);
      }
    }
  }
  __syncthreads();
  const int nf = (kF - f0) < 32 ? (kF - f0) : 32;
  for (int item = threadIdx.x; item < nf * 41; item += 256) {
    const int c = item / 41, tb = item - c * 41;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(tile + c * kXpPitch + 8 * tb);
    uint16_t* dst = xt + ((long long)lead_rows + ((long long)n * cols + f0 + c + col_pad) * 41 + tb) * 8;
    st_global_v4(dst, p[0], p[1], p[2], p[3]);
    if constexpr (SPLIT) {
      const uint32_t* pl = reinterpret_cast<const uint32_t*>(tile + kLo + c * kXpPitch + 8 * tb);
      st_global_v4(xt_lo + (dst - xt), pl[0], pl[1], pl[2], pl[3]);
    }
  }
}

}  // namespace
}  // namespace dfs
