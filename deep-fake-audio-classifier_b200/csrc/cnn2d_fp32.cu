// cnn2d_fp32.cu -- the 2D-CNN (/root/reference/src/model.py:14-42) in full fp32 on the CUDA cores: option "precision" = 1.
//
// Why it exists.  The tcgen05 path rounds the features, the folded weights and both intermediate activations to fp16
// (fp32 accumulation): scores agree with the reference to ~2e-5 relative, far inside the 1e-3 score gate, but with
// random-init weights neighbouring scores lie ~4e-6 apart, so that error re-orders neighbours and the EER of a few
// thousand utterances moves by whole quanta of 1/(2 n_class).  The reference is self-consistent to ~2e-7 across thread
// counts / batch sizes (tools/experiments/reference_self_consistency.py); this path has the same arithmetic class --
// fp32 operands, fp32 FMA accumulation, BN folded in double and rounded once -- and exists for evaluations where the rank
// order of near-equal scores matters (dev-set EER of a few thousand utterances).  It is an explicit option, never a
// fallback: ~35x slower than the tensor-core path and still ~60x the reference's 16-core CPU loop.  The FMAs are packed
// FFMA2 (fma.rn.f32x2, activation broadcast): 288 instead of 576 FMA instructions per channel step, 7.4 k -> 8.4 k utt/s.
//
// Layout: activations channels-last fp32 [n][rows][180][C]; weights [(kh*3+kw)*CI + ci][co] fp32.
// One thread = 2 rows x 2 feature columns x 4 output channels (the 2 rows are the pair the (2,1) average pool combines):
// 16 input vectors (4 rows x 4 columns, 4 channels each) are loaded once per channel step and reused by 36 weight vectors.
#include "common.cuh"
#include "kernels.h"

namespace dfs {

// in (FIRST): raw features, element (n,t,f) at in[n*sn + t*st + f*sf].  H = input rows, output rows = H/2 (POOL) or H.
template <int CI, int CO, bool FIRST, bool POOL>
__global__ void __launch_bounds__(128) conv3x3_fp32_kernel(const float* __restrict__ in, long long sn, long long st, long long sf, int H,
                                                           const float* __restrict__ w, const float* __restrict__ b, long long total,
                                                           float* __restrict__ out) {
  constexpr int CG = CO / 4, FP = kF / 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co = 4 * (int)(idx % CG);
  long long p = idx / CG;
  const int f0 = 2 * (int)(p % FP); p /= FP;
  const int HP = H / 2;
  const int j = (int)(p % HP);
  const long long n = p / HP;

  // accumulators as packed fp32 pairs: one FFMA2 (fma.rn.f32x2) = two IEEE fp32 FMAs with the activation broadcast
  uint64_t acc[2][2][2];   // [row of the pair][column of the pair][channel pair]
  {
    const float4 bv = *reinterpret_cast<const float4*>(b + co);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) { acc[a][c][0] = pack_f32x2(bv.x, bv.y); acc[a][c][1] = pack_f32x2(bv.z, bv.w); }
  }
  constexpr int CS = FIRST ? 1 : 4;   // channels per step
  for (int ci = 0; ci < CI; ci += CS) {
    float v[4][4][CS];   // rows 2j-1 .. 2j+2, columns f0-1 .. f0+2
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = 2 * j - 1 + r;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = f0 - 1 + c;
        const bool ok = row >= 0 && row < H && col >= 0 && col < kF;
        if constexpr (FIRST) {
          v[r][c][0] = ok ? in[n * sn + (long long)row * st + (long long)col * sf] : 0.0f;
        } else {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) t = *reinterpret_cast<const float4*>(in + ((n * H + row) * kF + col) * CI + ci);
          v[r][c][0] = t.x; v[r][c][1] = t.y; v[r][c][2] = t.z; v[r][c][3] = t.w;
        }
      }
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int q = 0; q < CS; ++q) {
          const float4 wv = *reinterpret_cast<const float4*>(w + ((long long)((kh * 3 + kw) * CI + ci + q)) * CO + co);
          const uint64_t w01 = pack_f32x2(wv.x, wv.y), w23 = pack_f32x2(wv.z, wv.w);
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const float x = v[a + kh][c + kw][q];
              const uint64_t xx = pack_f32x2(x, x);
              acc[a][c][0] = fma_f32x2(xx, w01, acc[a][c][0]);
              acc[a][c][1] = fma_f32x2(xx, w23, acc[a][c][1]);
            }
        }
  }
  float accf[2][2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      unpack_f32x2(acc[a][c][0], accf[a][c][0], accf[a][c][1]);
      unpack_f32x2(acc[a][c][1], accf[a][c][2], accf[a][c][3]);
    }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if constexpr (POOL) {   // relu, then the (2,1) average (model.py:18,24)
      float4 o;
      o.x = 0.5f * (relu_nan(accf[0][c][0]) + relu_nan(accf[1][c][0]));
      o.y = 0.5f * (relu_nan(accf[0][c][1]) + relu_nan(accf[1][c][1]));
      o.z = 0.5f * (relu_nan(accf[0][c][2]) + relu_nan(accf[1][c][2]));
      o.w = 0.5f * (relu_nan(accf[0][c][3]) + relu_nan(accf[1][c][3]));
      *reinterpret_cast<float4*>(out + ((n * HP + j) * kF + f0 + c) * CO + co) = o;
    } else {
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const float4 o = make_float4(relu_nan(accf[a][c][0]), relu_nan(accf[a][c][1]), relu_nan(accf[a][c][2]), relu_nan(accf[a][c][3]));
        *reinterpret_cast<float4*>(out + ((n * H + 2 * j + a) * kF + f0 + c) * CO + co) = o;
      }
    }
  }
}

// emb[n][f][c] = sum_t act3[n][t][f][c] (fixed order; the head's weights carry 1/80)
__global__ void __launch_bounds__(256) time_sum_fp32_kernel(const float* __restrict__ act3, long long total, float* __restrict__ emb) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  constexpr long long PER = (long long)kF * 128;
  const long long n = idx / PER, r = idx - n * PER;
  const float* src = act3 + n * 80 * PER + r;
  double s = 0.0;
#pragma unroll 8
  for (int t = 0; t < 80; ++t) s += (double)src[t * PER];
  emb[idx] = (float)s;
}

// logits with an fp64 dot (23,040 terms): one block per utterance
__global__ void __launch_bounds__(256) cnn2d_head_f64_kernel(const float* __restrict__ emb, const float* __restrict__ wfc, float fcb,
                                                             int apply_sigmoid, float* __restrict__ out) {
  constexpr int NE = kF * 128;
  const float* e = emb + (long long)blockIdx.x * NE;
  double acc = 0.0;
  for (int i = threadIdx.x; i < NE; i += blockDim.x) acc = fma((double)e[i], (double)wfc[i], acc);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i];
    const float z = (float)(s + (double)fcb);
    out[blockIdx.x] = apply_sigmoid ? 1.0f / (1.0f + expf(-z)) : z;
  }
}

size_t cnn2d_fp32_work_floats(int n_utts) { return (size_t)n_utts * kF * (160 * 32 + 80 * 64 + 80 * 128); }

int launch_cnn2d_fp32(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, const float* const w[3], const float* const b[3],
                      const float* wfc, float fcb, int apply_sigmoid, float* work, float* emb, float* out, cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  float* a1 = work;
  float* a2 = a1 + (size_t)n_utts * 160 * kF * 32;
  float* a3 = a2 + (size_t)n_utts * 80 * kF * 64;
  long long tot = (long long)n_utts * 160 * (kF / 2) * (32 / 4);
  conv3x3_fp32_kernel<1, 32, true, true><<<(unsigned)ceil_div64(tot, 128), 128, 0, stream>>>(x, sn, st, sf, kT, w[0], b[0], tot, a1);
  DFS_LAUNCH_CHECK();
  tot = (long long)n_utts * 80 * (kF / 2) * (64 / 4);
  conv3x3_fp32_kernel<32, 64, false, true><<<(unsigned)ceil_div64(tot, 128), 128, 0, stream>>>(a1, 0, 0, 0, 160, w[1], b[1], tot, a2);
  DFS_LAUNCH_CHECK();
  tot = (long long)n_utts * 40 * (kF / 2) * (128 / 4);
  conv3x3_fp32_kernel<64, 128, false, false><<<(unsigned)ceil_div64(tot, 128), 128, 0, stream>>>(a2, 0, 0, 0, 80, w[2], b[2], tot, a3);
  DFS_LAUNCH_CHECK();
  tot = (long long)n_utts * kF * 128;
  time_sum_fp32_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, stream>>>(a3, tot, emb);
  DFS_LAUNCH_CHECK();
  cnn2d_head_f64_kernel<<<n_utts, 256, 0, stream>>>(emb, wfc, fcb, apply_sigmoid, out);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

}  // namespace dfs
