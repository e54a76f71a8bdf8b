// simt_models.cu -- first-correct CUDA-core (fp32) kernels for the two scorers that are not yet
// on the tensor-core path:
//   CNN1D            /root/reference/src/model_cnn1d.py:37-46  (3 x Conv1d k3 + BN + ReLU, mean over T, Linear)
//   ConvAutoencoder  /root/reference/src/model_cae.py:83-125   (4 x conv/BN/ReLU/AvgPool2d(2), 3 x ConvT k2s2/BN/ReLU,
//                    ConvT 32->1) fused with the per-utterance MSE of src/predict_hybrid.py:75-76.
// All activations are channels-last fp32; BN is folded into the weights at create time (api.cu).
// These kernels are plain, coalesced and correct; DESIGN.md lists them as the next to move onto
// the tcgen05 template of conv_tc.cu.
#include "common.cuh"
#include "kernels.h"

namespace dfs {

// ------------------------------------------------------------------------------------------
// CNN1D
// ------------------------------------------------------------------------------------------
// out[n][t][co] = relu(b[co] + sum_k sum_ci w[(k*CI + ci)*CO + co] * in(n, t+k-1, ci))
// IN_STRIDED: `in` is the raw feature tensor addressed through (sn, st, sf); else channels-last [n][321][CI].
template <bool IN_STRIDED>
__global__ void conv1d_k3_kernel(const float* __restrict__ in, long long sn, long long st, long long sf, int CI, int CO,
                                 const float* __restrict__ w, const float* __restrict__ b, long long total, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co = (int)(idx % CO);
  const long long pos = idx / CO;
  const int t = (int)(pos % kT);
  const long long n = pos / kT;
  float acc = b[co];
  for (int k = 0; k < 3; ++k) {
    const int tt = t + k - 1;
    if (tt < 0 || tt >= kT) continue;
    const float* src = IN_STRIDED ? in + n * sn + tt * st : in + (n * kT + tt) * CI;
    const long long cs = IN_STRIDED ? sf : 1;
    const float* wk = w + (long long)k * CI * CO + co;
    for (int ci = 0; ci < CI; ++ci) acc = fmaf(src[ci * cs], wk[(long long)ci * CO], acc);
  }
  out[idx] = relu_nan(acc);
}

// logits[n] = fcb + sum_c fcw[c] * mean_t h[n][t][c]; one block (128 threads = channels) per utterance
__global__ void __launch_bounds__(128) cnn1d_head_kernel(const float* __restrict__ h, const float* __restrict__ fcw, float fcb,
                                                          int apply_sigmoid, float* __restrict__ out) {
  const int c = threadIdx.x;
  const float* src = h + (long long)blockIdx.x * kT * 128 + c;
  float s = 0.0f;
  for (int t = 0; t < kT; ++t) s += src[t * 128];
  float v = (s / (float)kT) * fcw[c];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ float part[4];
  if ((c & 31) == 0) part[c >> 5] = v;
  __syncthreads();
  if (c == 0) {
    const float z = fcb + part[0] + part[1] + part[2] + part[3];
    out[blockIdx.x] = apply_sigmoid ? 1.0f / (1.0f + expf(-z)) : z;
  }
}

size_t cnn1d_simt_work_floats(int n_utts) { return (size_t)n_utts * kT * (32 + 64 + 128); }

int launch_cnn1d_simt(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, const SimtConv* cv, const float* fcw, float fcb,
                      int apply_sigmoid, float* work, float* out, cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  float* h1 = work;
  float* h2 = h1 + (size_t)n_utts * kT * 32;
  float* h3 = h2 + (size_t)n_utts * kT * 64;
  long long tot = (long long)n_utts * kT * 32;
  conv1d_k3_kernel<true><<<(unsigned)ceil_div64(tot, 256), 256, 0, stream>>>(x, sn, st, sf, kF, 32, cv[0].w, cv[0].b, tot, h1);
  DFS_LAUNCH_CHECK();
  tot = (long long)n_utts * kT * 64;
  conv1d_k3_kernel<false><<<(unsigned)ceil_div64(tot, 256), 256, 0, stream>>>(h1, 0, 0, 0, 32, 64, cv[1].w, cv[1].b, tot, h2);
  DFS_LAUNCH_CHECK();
  tot = (long long)n_utts * kT * 128;
  conv1d_k3_kernel<false><<<(unsigned)ceil_div64(tot, 256), 256, 0, stream>>>(h2, 0, 0, 0, 64, 128, cv[2].w, cv[2].b, tot, h3);
  DFS_LAUNCH_CHECK();
  cnn1d_head_kernel<<<n_utts, 128, 0, stream>>>(h3, fcw, fcb, apply_sigmoid, out);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

// ------------------------------------------------------------------------------------------
// CAE
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float cae_input(const float* __restrict__ x, long long sn, long long st, long long sf, long long n, int t,
                                           int f, const float* __restrict__ mean, const float* __restrict__ sd) {
  float v = x[n * sn + t * st + f * sf];
  if (mean != nullptr) v = (v - mean[f]) / sd[f];
  return v;
}

// encoder block: conv3x3(p=1) + folded BN + ReLU + AvgPool2d(2) (floor).  in: channels-last [n][H][W][CI]
// (or the strided raw input when FIRST), out: [n][H/2][W/2][CO].  w: [(kh*3+kw)*CI + ci][co].
template <bool FIRST>
__global__ void cae_enc_kernel(const float* __restrict__ in, long long sn, long long st, long long sf, const float* __restrict__ mean,
                               const float* __restrict__ sd, int H, int W, int CI, int CO, const float* __restrict__ w,
                               const float* __restrict__ b, long long total, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Ho = H / 2, Wo = W / 2;
  const int co = (int)(idx % CO);
  long long pos = idx / CO;
  const int wo = (int)(pos % Wo); pos /= Wo;
  const int ho = (int)(pos % Ho);
  const long long n = pos / Ho;
  float pooled = 0.0f;
  for (int dy = 0; dy < 2; ++dy)
    for (int dx = 0; dx < 2; ++dx) {
      const int y = 2 * ho + dy, xq = 2 * wo + dx;
      float acc = b[co];
      for (int kh = 0; kh < 3; ++kh) {
        const int yy = y + kh - 1;
        if (yy < 0 || yy >= H) continue;
        for (int kw = 0; kw < 3; ++kw) {
          const int xx = xq + kw - 1;
          if (xx < 0 || xx >= W) continue;
          const float* wk = w + (long long)((kh * 3 + kw) * CI) * CO + co;
          if (FIRST) {
            acc = fmaf(cae_input(in, sn, st, sf, n, yy, xx, mean, sd), wk[0], acc);
          } else {
            const float* src = in + ((n * H + yy) * W + xx) * CI;
            for (int ci = 0; ci < CI; ++ci) acc = fmaf(src[ci], wk[(long long)ci * CO], acc);
          }
        }
      }
      pooled += relu_nan(acc);
    }
  out[idx] = pooled * 0.25f;
}

// decoder block: ConvTranspose2d(k=2,s=2[,output_padding=(0,opw)]) + folded BN + ReLU.
// in [n][H][W][CI] -> out [n][2H][2W+opw][CO].  w: [((a*2+b)*CI + ci)][co] (BN scale folded in).
__global__ void cae_dec_kernel(const float* __restrict__ in, int H, int W, int CI, int CO, int opw, const float* __restrict__ w,
                               const float* __restrict__ b, long long total, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Ho = 2 * H, Wo = 2 * W + opw;
  const int co = (int)(idx % CO);
  long long pos = idx / CO;
  const int xo = (int)(pos % Wo); pos /= Wo;
  const int yo = (int)(pos % Ho);
  const long long n = pos / Ho;
  float acc = b[co];
  if (xo < 2 * W) {  // the output_padding column receives the bias only
    const int i = yo >> 1, a = yo & 1, j = xo >> 1, bb = xo & 1;
    const float* src = in + ((n * H + i) * W + j) * CI;
    const float* wk = w + (long long)((a * 2 + bb) * CI) * CO + co;
    for (int ci = 0; ci < CI; ++ci) acc = fmaf(src[ci], wk[(long long)ci * CO], acc);
  }
  out[idx] = relu_nan(acc);
}

// final ConvTranspose2d(32,1,k2,s2) (no BN / activation), zero row 320, fused per-utterance MSE:
// one block per utterance, fixed summation order.  w: [(a*2+b)*32 + ci].
__global__ void __launch_bounds__(256) cae_final_mse_kernel(const float* __restrict__ d3 /*[n][160][90][32]*/, const float* __restrict__ x,
                                                             long long sn, long long st, long long sf, const float* __restrict__ mean,
                                                             const float* __restrict__ sd, const float* __restrict__ w, float bias,
                                                             float* __restrict__ mse_out, float* __restrict__ recon_out) {
  const long long n = blockIdx.x;
  __shared__ float ws[4 * 32];
  if (threadIdx.x < 128) ws[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  float acc = 0.0f;
  for (int e = threadIdx.x; e < kT * kF; e += blockDim.x) {
    const int t = e / kF, f = e - t * kF;
    float r = 0.0f;
    if (t < 320) {
      const float* src = d3 + ((n * 160 + (t >> 1)) * 90 + (f >> 1)) * 32;
      const float* wk = ws + ((t & 1) * 2 + (f & 1)) * 32;
      r = bias;
#pragma unroll 8
      for (int ci = 0; ci < 32; ++ci) r = fmaf(src[ci], wk[ci], r);
    }
    if (recon_out != nullptr) recon_out[n * kT * kF + e] = r;
    const float d = r - cae_input(x, sn, st, sf, n, t, f, mean, sd);
    acc = fmaf(d, d, acc);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && mse_out != nullptr) {
    float s = 0.0f;
    for (int i = 0; i < 8; ++i) s += part[i];
    mse_out[n] = s / (float)(kT * kF);
  }
}

__global__ void cae_latent_export_kernel(const float* __restrict__ e4 /*[n][20][11][256]*/, long long total, float* __restrict__ latent) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // latent index [n][256][20][11]
  if (idx >= total) return;
  const int xw = (int)(idx % 11);
  const int y = (int)((idx / 11) % 20);
  const int c = (int)((idx / 220) % 256);
  const long long n = idx / (220 * 256);
  latent[idx] = e4[((n * 20 + y) * 11 + xw) * 256 + c];
}

// per-utterance fp32 work: e1 160x90x32, e2 80x45x64, e3 40x22x128, e4 20x11x256, d1 40x22x128, d2 80x45x64, d3 160x90x32
static constexpr size_t kCaeWork[7] = {160 * 90 * 32, 80 * 45 * 64, 40 * 22 * 128, 20 * 11 * 256, 40 * 22 * 128, 80 * 45 * 64, 160 * 90 * 32};
size_t cae_simt_work_floats(int n_utts) {
  size_t s = 0;
  for (size_t v : kCaeWork) s += v;
  return s * (size_t)n_utts;
}

const float* cae_simt_layer_ptr(const float* work, int n_utts, int layer, size_t* floats_per_utt) {
  const float* p = work;
  for (int i = 0; i < layer; ++i) p += kCaeWork[i] * (size_t)n_utts;
  *floats_per_utt = kCaeWork[layer];
  return p;
}

int launch_cae_simt(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, const SimtConv* enc, const SimtConv* dec,
                    float final_bias, const float* norm_mean, const float* norm_std, float* work, float* mse_out, float* recon_out, float* latent_out,
                    cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  float* buf[7];
  float* p = work;
  for (int i = 0; i < 7; ++i) { buf[i] = p; p += kCaeWork[i] * (size_t)n_utts; }
  const int B = 128;
  auto blocks = [&](long long tot) { return (unsigned)ceil_div64(tot, B); };
  long long tot = (long long)kCaeWork[0] * n_utts;
  cae_enc_kernel<true><<<blocks(tot), B, 0, stream>>>(x, sn, st, sf, norm_mean, norm_std, kT, kF, 1, 32, enc[0].w, enc[0].b, tot, buf[0]);
  DFS_LAUNCH_CHECK();
  tot = (long long)kCaeWork[1] * n_utts;
  cae_enc_kernel<false><<<blocks(tot), B, 0, stream>>>(buf[0], 0, 0, 0, nullptr, nullptr, 160, 90, 32, 64, enc[1].w, enc[1].b, tot, buf[1]);
  DFS_LAUNCH_CHECK();
  tot = (long long)kCaeWork[2] * n_utts;
  cae_enc_kernel<false><<<blocks(tot), B, 0, stream>>>(buf[1], 0, 0, 0, nullptr, nullptr, 80, 45, 64, 128, enc[2].w, enc[2].b, tot, buf[2]);
  DFS_LAUNCH_CHECK();
  tot = (long long)kCaeWork[3] * n_utts;
  cae_enc_kernel<false><<<blocks(tot), B, 0, stream>>>(buf[2], 0, 0, 0, nullptr, nullptr, 40, 22, 128, 256, enc[3].w, enc[3].b, tot, buf[3]);
  DFS_LAUNCH_CHECK();
  if (latent_out != nullptr) {
    cae_latent_export_kernel<<<blocks(tot), B, 0, stream>>>(buf[3], tot, latent_out);
    DFS_LAUNCH_CHECK();
  }
  tot = (long long)kCaeWork[4] * n_utts;
  cae_dec_kernel<<<blocks(tot), B, 0, stream>>>(buf[3], 20, 11, 256, 128, 0, dec[0].w, dec[0].b, tot, buf[4]);
  DFS_LAUNCH_CHECK();
  tot = (long long)kCaeWork[5] * n_utts;
  cae_dec_kernel<<<blocks(tot), B, 0, stream>>>(buf[4], 40, 22, 128, 64, 1, dec[1].w, dec[1].b, tot, buf[5]);
  DFS_LAUNCH_CHECK();
  tot = (long long)kCaeWork[6] * n_utts;
  cae_dec_kernel<<<blocks(tot), B, 0, stream>>>(buf[5], 80, 45, 64, 32, 0, dec[2].w, dec[2].b, tot, buf[6]);
  DFS_LAUNCH_CHECK();
  cae_final_mse_kernel<<<n_utts, 256, 0, stream>>>(buf[6], x, sn, st, sf, norm_mean, norm_std, dec[3].w, final_bias, mse_out, recon_out);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

}  // namespace dfs
