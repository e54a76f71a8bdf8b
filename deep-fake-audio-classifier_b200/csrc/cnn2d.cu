// cnn2d.cu -- the CUDA-core stages around the tensor-core convolutions of the 2D-CNN scorer
// (/root/reference/src/model.py:12-42):
//   conv1_kernel      Conv2d(1,32,3,p=1)+BN+ReLU+AvgPool2d((2,1))  (model.py:15-18) -> FT8 fp16
//                     (Cin = 1, K = 9 is not a tensor-core shape; also the fp32 -> fp16 / layout producer)
//                     with POOLF it is the CAE encoder block 1 (model_cae.py:34-37, AvgPool2d(2)).
//   head_kernel       x.mean(dim=2) -> flatten -> Linear(23040,1) [-> sigmoid]  (model.py:37-39, predict.py:108)
//   *_simt kernels    conv2 / conv3 on CUDA cores over the SAME packed weights and layouts: a debug
//                     cross-check for the tcgen05 kernels (option conv_impl = 1), never the default.
#include "common.cuh"
#include "kernels.h"
#include "layout.cuh"

namespace dfs {

// ------------------------------------------------------------------------------------------
// conv1: one block = one utterance x one chunk of 15 output feature columns, all 160 pooled rows
// ------------------------------------------------------------------------------------------
template <bool POOLF>
__global__ void __launch_bounds__(256) conv1_kernel(const float* __restrict__ x, long long sn, long long st, long long sf,
                                                     const __grid_constant__ Conv1Weights w, const float* __restrict__ norm_mean,
                                                     const float* __restrict__ norm_std, ActBuf out, int cols_out) {
  constexpr int FO = 15;                    // output feature columns per block
  constexpr int FW = POOLF ? 30 : 15;       // input feature columns per block (without halo)
  constexpr int XW = POOLF ? 33 : 17;       // smem row pitch (FW + 2 halo, padded against bank conflicts)
  constexpr int NC = POOLF ? 4 : 3;         // input columns feeding one output
  constexpr int TR = kT + 2;                // rows incl. zero halo
  __shared__ float xs[TR * XW];
  const int n = blockIdx.y;
  const int f0 = blockIdx.x * FW;
  const float* xn = x + (long long)n * sn;
  for (int idx = threadIdx.x; idx < TR * (FW + 2); idx += blockDim.x) {
    int tr, fc;
    if (sf == 1) { tr = idx / (FW + 2); fc = idx - tr * (FW + 2); }   // feature-contiguous storage
    else         { fc = idx / TR;       tr = idx - fc * TR; }         // time-contiguous storage (reference's transposed view)
    const int t = tr - 1, f = f0 - 1 + fc;
    float v = 0.0f;
    if (t >= 0 && t < kT && f >= 0 && f < kF) {
      v = xn[t * st + f * sf];
      if (norm_mean != nullptr) v = (v - norm_mean[f]) / norm_std[f];  // dataset_cae.py:41
    }
    xs[tr * XW + fc] = v;
  }
  __syncthreads();

  const long long plane_elems = out.plane_elems();
  for (int o = threadIdx.x; o < FO * 160; o += blockDim.x) {
    const int fl = o / 160, j = o - fl * 160;
    float xin[4][NC];
    const int fc0 = POOLF ? 2 * fl : fl;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < NC; ++b) xin[a][b] = xs[(2 * j + a) * XW + fc0 + b];
    const int fo = blockIdx.x * FO + fl;
    const long long gc = (long long)n * cols_out + fo + 1;
    // FT8P (layout.cuh): pooled time step j -> parity plane group j&1, row j/2 + 1 (both consumers are PAIR GEMMs)
    uint16_t* dst = out.ptr + (long long)((j & 1) * 4) * plane_elems + (gc * out.RS + (j >> 1) + 1) * 8;
#pragma unroll
    for (int pj = 0; pj < 4; ++pj) {
      float r[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = pj * 8 + e;
        float acc = 0.0f;
#pragma unroll
        for (int dt = 0; dt < 2; ++dt)
#pragma unroll
          for (int df = 0; df < (POOLF ? 2 : 1); ++df) {
            float a = w.b[c];
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) a = fmaf(w.w[c * 9 + kh * 3 + kw], xin[kh + dt][kw + df], a);
            acc += relu_nan(a);
          }
        r[e] = acc * (POOLF ? 0.25f : 0.5f);
      }
      st_global_v4(dst + pj * plane_elems, pack_act2(r[0], r[1]), pack_act2(r[2], r[3]), pack_act2(r[4], r[5]),
                   pack_act2(r[6], r[7]));
    }
  }
}

int launch_conv1(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, const Conv1Weights& w, const float* norm_mean,
                 const float* norm_std, bool pool_f, ActBuf out, cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  if (pool_f) {
    conv1_kernel<true><<<dim3(6, n_utts), 256, 0, stream>>>(x, sn, st, sf, w, norm_mean, norm_std, out, kF / 2 + 2);
  } else {
    conv1_kernel<false><<<dim3(12, n_utts), 256, 0, stream>>>(x, sn, st, sf, w, norm_mean, norm_std, out, kCols);
  }
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

// ------------------------------------------------------------------------------------------
// head: one block per utterance
// ------------------------------------------------------------------------------------------
// ACC = double ("split" precision): the 23,040 products are exact in fp64 and so is their sum to ~1e-16; with trained-like
// classifiers the positive and negative terms cancel to ~1e-3 of their absolute sum, and an fp32 accumulation order shows up
// as 1e-4 of logit -- more than the three split-precision conv layers together.
template <typename ACC>
__global__ void __launch_bounds__(256) cnn2d_head_kernel(const float* __restrict__ emb, const float* __restrict__ wfc, float fcb,
                                                          int apply_sigmoid, float* __restrict__ out) {
  constexpr int NE = kF * 128;
  const float4* e4 = reinterpret_cast<const float4*>(emb + (long long)blockIdx.x * NE);
  const float4* w4 = reinterpret_cast<const float4*>(wfc);
  ACC acc = 0;
  for (int i = threadIdx.x; i < NE / 4; i += blockDim.x) {
    const float4 a = e4[i], b = w4[i];
    if constexpr (sizeof(ACC) == 8) {
      acc = fma((double)a.x, (double)b.x, acc); acc = fma((double)a.y, (double)b.y, acc);
      acc = fma((double)a.z, (double)b.z, acc); acc = fma((double)a.w, (double)b.w, acc);
    } else {
      acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ ACC part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    ACC s = (ACC)fcb;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i];
    const float sf = (float)s;
    out[blockIdx.x] = apply_sigmoid ? 1.0f / (1.0f + expf(-sf)) : sf;
  }
}

int launch_cnn2d_head(const float* emb, const float* wfc, float fcb, int n_utts, int apply_sigmoid, float* out, cudaStream_t stream, bool acc64) {
  if (n_utts <= 0) return DFS_OK;
  if (acc64) cnn2d_head_kernel<double><<<n_utts, 256, 0, stream>>>(emb, wfc, fcb, apply_sigmoid, out);
  else cnn2d_head_kernel<float><<<n_utts, 256, 0, stream>>>(emb, wfc, fcb, apply_sigmoid, out);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

__global__ void cnn2d_embedding_export_kernel(const float* __restrict__ emb, float* __restrict__ embedding, long long total) {
  // out index = n*23040 + c*180 + f  <-  emb[n][f][c] / 80
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int f = (int)(idx % kF);
  const int c = (int)((idx / kF) % 128);
  const long long n = idx / (kF * 128);
  embedding[idx] = emb[(n * kF + f) * 128 + c] / 80.0f;
}

int launch_cnn2d_embedding_export(const float* emb, int n_utts, float* embedding, cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  const long long total = (long long)n_utts * kF * 128;
  cnn2d_embedding_export_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, stream>>>(emb, embedding, total);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

// ------------------------------------------------------------------------------------------
// debug cross-check kernels (CUDA cores, same data as the tcgen05 kernels)
// ------------------------------------------------------------------------------------------
template <int CIN, int COUT>
__device__ __forceinline__ float conv_at(const ActBuf& a, const uint16_t* __restrict__ wpack, long long gc, int tp, int co) {
  // padded position (gc, tp); taps reach gc-1..gc+1, tp-1..tp+1 (zero pads make the borders right)
  float acc = 0.0f;
  const long long plane_elems = a.plane_elems();
  for (int tap = 0; tap < 9; ++tap) {
    const int kh = tap / 3, kw = tap % 3;
    const uint16_t* src = a.ptr + ((gc + kw - 1) * a.RS + (tp + kh - 1)) * 8;
    for (int ci = 0; ci < CIN; ++ci) {
      const float xv = act_bits_to_float(src[(ci >> 3) * plane_elems + (ci & 7)]);
      const float wv = act_bits_to_float(wpack[(((long long)tap * (CIN / 8) + (ci >> 3)) * COUT + co) * 8 + (ci & 7)]);
      acc = fmaf(xv, wv, acc);
    }
  }
  return acc;
}

__global__ void cnn2d_conv2_simt_kernel(ActBuf act1, const uint16_t* __restrict__ wpack, const float* __restrict__ bias, long long total,
                                        ActBuf act2) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co = (int)(idx % 64);
  const long long pos = idx / 64;
  const int to = (int)(pos % 80);
  const int f = (int)((pos / 80) % kF);
  const long long n = pos / (80 * kF);
  const long long gc = n * kCols + f + 1;
  // Same GEMM as the PAIR formulation of conv_tc.cu, evaluated naively: output pair `to`, column dt2*64 + co,
  // K = (input time step r of 2*to-1 .. 2*to+2, feature tap kw, channel); act1 is FT8P (row = t/2 + 1, parity planes);
  // wpack = [r*3+kw][ci/8][128][8] with the 0.5 of the average pool folded in (bias too).
  const long long plane_elems = act1.plane_elems();
  float acc[2] = {0.0f, 0.0f};
  for (int r = 0; r < 4; ++r) {
    const int t_in = 2 * to - 1 + r;                       // -1 .. 160; s = t_in + 2 = 2*row + par
    const int par = (t_in + 2) & 1, row = (t_in + 2) >> 1;
    for (int kw = 0; kw < 3; ++kw) {
      const uint16_t* src = act1.ptr + (long long)(par * 4) * plane_elems + ((gc + kw - 1) * act1.RS + row) * 8;
      for (int ci = 0; ci < 32; ++ci) {
        const float xv = act_bits_to_float(src[(ci >> 3) * plane_elems + (ci & 7)]);
        const uint16_t* wrow = wpack + ((((long long)(r * 3 + kw)) * 4 + (ci >> 3)) * 128) * 8 + (ci & 7);
        acc[0] = fmaf(xv, act_bits_to_float(wrow[(long long)co * 8]), acc[0]);
        acc[1] = fmaf(xv, act_bits_to_float(wrow[(long long)(64 + co) * 8]), acc[1]);
      }
    }
  }
  const float s = relu_nan(acc[0] + bias[co]) + relu_nan(acc[1] + bias[co]);
  const __half b = __float2half_rn(fminf(s, 65504.0f));
  act2.ptr[(co >> 3) * act2.plane_elems() + (gc * act2.RS + to + 1) * 8 + (co & 7)] = *reinterpret_cast<const uint16_t*>(&b);
}

__global__ void cnn2d_conv3_simt_kernel(ActBuf act2, const uint16_t* __restrict__ wpack, const float* __restrict__ bias, long long total,
                                        float* __restrict__ emb) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co = (int)(idx % 128);
  const long long pos = idx / 128;
  const int f = (int)(pos % kF);
  const long long n = pos / kF;
  const long long gc = n * kCols + f + 1;
  float s = 0.0f;
  for (int tp = 1; tp <= 80; ++tp) s += relu_nan(conv_at<64, 128>(act2, wpack, gc, tp, co) + bias[co]);
  emb[idx] = s;  // idx == (n*180 + f)*128 + co
}

int launch_cnn2d_conv2_simt(ActBuf act1, const uint16_t* wpack, const float* bias_dev, int n_utts, ActBuf act2, cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  const long long total = (long long)n_utts * kF * 80 * 64;
  cnn2d_conv2_simt_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, stream>>>(act1, wpack, bias_dev, total, act2);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

int launch_cnn2d_conv3_simt(ActBuf act2, const uint16_t* wpack, const float* bias_dev, int n_utts, float* emb, cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  const long long total = (long long)n_utts * kF * 128;
  cnn2d_conv3_simt_kernel<<<(unsigned)ceil_div64(total, 128), 128, 0, stream>>>(act2, wpack, bias_dev, total, emb);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

}  // namespace dfs
