// cae_tc.cu -- the convolutional autoencoder scorer on the tcgen05 template (conv_tc.cuh):
//   encoder  4 x [Conv2d 3x3 p=1 + BN + ReLU + AvgPool2d(2)]          /root/reference/src/model_cae.py:32-56
//   decoder  3 x [ConvTranspose2d k=2 s=2 + BN + ReLU] + ConvT 32->1  /root/reference/src/model_cae.py:61-81
//   score    MSELoss(reduction='none')(recon, x).view(B,-1).mean(1)   /root/reference/src/predict_hybrid.py:75-76
//
//   layer  in (C x T x F)      kernel                                              out layout (planes, cols, RS)
//   enc1   1 x 321 x 180       Toeplitz-in-time GEMM N=256 (cae_enc1_tc.cu; normaliser in the prep) e1 FT8P ( 8, 92,  82)
//   enc2   32 x 160 x 90       PAIR GEMM  N=128, K=384, time pool in-thread + lane^8 e2 FT8  ( 8, 48,  82)
//   enc3   64 x 80 x 45        3x3 GEMM   swapped roles: weights = A, N=256 positions, K=576, 2x2 pool in-thread   e3 FT8  (16, 24,  42)
//          (option "enc3_swap" = 0: positions as M, N=128, pool with lane^1 / lane^8; scores agree to 1e-5)
//   enc4   128 x 40 x 22       3x3 GEMM   4 groups of N=64, K=1152 in 2 pieces       e4 FT8  (32, 14,  26)   = latent
//   dec1   256 x 20 x 11       1x1 GEMM   2 groups (b) of N=256=(32-ch block, a, 32), K=256    d1 FT8  (16, 24,  42)
//   dec2   128 x 40 x 22       1x1 GEMM   1 group of N=256=(b, 32-ch block, a, 32), K=128      d2 FT8  ( 8, 48,  82)   col 45 = relu(bias)
//          (option "dec_wide" = 0: the N = 128 variants with 4 / 2 groups, bit-identical, 1.3-1.5x slower: each group re-reads the input)
//   dec3   64 x 80 x 45        1x1 GEMM   N=(a,b,32), K=64                           d3 FT8  ( 4, 92, 162)
//   final  32 x 160 x 90       fused into dec3's epilogue on the scoring path (EPI_SHUFFLE_MSE: 4 outputs x 32 MACs per d3 vector,
//                              residual vs the (normalised) input, one partial sum per 16-column unit; neither d3 nor the
//                              reconstruction is written) + cae_mse_finish_kernel (zero row 320, per-utterance mean).
//                              cae_final_tc_kernel (reads d3) serves forward()'s materialised reconstruction and the cross-check.
// AvgPool2d(2) floors: enc1 drops input row 320, enc3 drops feature column 44 (out_feats = 22); dec2's
// output_padding column receives the bias only (constant, written once at handle creation).
#include <string.h>

#include "conv_tc.cuh"

namespace dfs {

using Enc2Cfg = ConvCfg<MODE_PAIR, 32, 64, 128, 80, 2, 3, 4, 1, EPI_PAIR_POOL_F>;
using Enc3Cfg = ConvCfg<MODE_3X3, 64, 128, 128, 80, 2, 2, 4, 1, EPI_POOL_TF>;
using Enc4Cfg = ConvCfg<MODE_3X3, 128, 64, 64, 40, 1, 3, 4, 2, EPI_POOL_TF>;
// enc3 with swapped operand roles (the GEMM shape of the 2D-CNN's conv3: weights = A, 256 positions = N, 83 % instead of 73 % of the
// tensor pipe per MMA); both pool partners are columns of one thread.  Option "enc3_swap".
using Enc3SwapCfg = ConvCfg<MODE_3X3S, 64, 128, 256, 80, 1, 3, 2, 2, EPI_POOL_TF_SWAP>;
// enc4's weights (147 KB per 64 output channels) force N = 64 on one CTA; a CTA pair runs N = 128 with 64 weight rows per CTA.
// enc2 / enc3 (already N = 128) were also tried on pairs: bit-identical and slower (CAE 351 k vs 362 k utt/s), see conv_tc.cu.
using Enc4PairCfg = ConvCfg<MODE_3X3, 128, 128, 128, 40, 1, 3, 4, 2, EPI_POOL_TF, 1>;   // option "pair_mma": CTA pairs, 2 groups of N = 128 (64 weight rows per CTA)
static_assert(Enc4PairCfg::PPL == Enc4Cfg::PPL && Enc4PairCfg::WROWS == Enc4Cfg::WROWS && Enc4PairCfg::WGT_B == Enc4Cfg::WGT_B, "enc4 pair variant shares map and weights");
using Dec1Cfg = ConvCfg<MODE_1X1, 256, 128, 128, 24, 1, 3, 2, 4, EPI_SHUFFLE_ROWS>;
using Dec2Cfg = ConvCfg<MODE_1X1, 128, 64, 128, 40, 1, 3, 2, 2, EPI_SHUFFLE_ROWS>;
using Dec1WideCfg = ConvCfg<MODE_1X1, 256, 128, 256, 24, 1, 5, 2, 4, EPI_SHUFFLE_ROWS>;   // option "dec_wide": N = 256, 2 groups
using Dec2WideCfg = ConvCfg<MODE_1X1, 128, 64, 256, 40, 1, 6, 2, 2, EPI_SHUFFLE_ROWS>;    //                    N = 256, 1 group
static_assert(Dec1WideCfg::PPL == Dec1Cfg::PPL && Dec2WideCfg::PPL == Dec2Cfg::PPL && Dec1WideCfg::WROWS == Dec1Cfg::WROWS, "the wide variants share the tensor maps");
using Dec3Cfg = ConvCfg<MODE_1X1, 64, 32, 128, 80, 2, 3, 4, 1, EPI_SHUFFLE>;
using Dec3MseCfg = ConvCfg<MODE_1X1, 64, 32, 128, 80, 2, 3, 2, 1, EPI_SHUFFLE_MSE>;   // dec3 + final ConvT + squared error, nothing written but partial sums

// geometry of the seven activation buffers: planes, padded cols per utterance, rows per column
static const int kCaePlanes[7] = {8, 8, 16, 32, 16, 8, 4};
static const int kCaeCols[7] = {92, 48, 24, 14, 24, 48, 92};
static const int kCaeRS[7] = {82, 82, 42, 26, 42, 82, 162};

void cae_tc_geometry(int layer, int* planes, int* cols, int* rs) {
  *planes = kCaePlanes[layer];
  *cols = kCaeCols[layer];
  *rs = kCaeRS[layer];
}

int cae_tc_make_maps(CaeTcState* s) {
  DFS_PROPAGATE(make_act_tensor_map(&s->tmap[0], s->act[0], Enc2Cfg::WROWS, Enc2Cfg::WCOLS, Enc2Cfg::PPL));
  DFS_PROPAGATE(make_act_tensor_map(&s->tmap[1], s->act[1], Enc3Cfg::WROWS, Enc3Cfg::WCOLS, Enc3Cfg::PPL));
  DFS_PROPAGATE(make_act_tensor_map(&s->tmap[2], s->act[2], Enc4Cfg::WROWS, Enc4Cfg::WCOLS, Enc4Cfg::PPL));
  DFS_PROPAGATE(make_act_tensor_map(&s->tmap_enc3_swap, s->act[1], Enc3SwapCfg::WROWS, Enc3SwapCfg::WCOLS, Enc3SwapCfg::PPL));
  DFS_PROPAGATE(make_act_tensor_map(&s->tmap[3], s->act[3], Dec1Cfg::WROWS, Dec1Cfg::WCOLS, Dec1Cfg::PPL));
  DFS_PROPAGATE(make_act_tensor_map(&s->tmap[4], s->act[4], Dec2Cfg::WROWS, Dec2Cfg::WCOLS, Dec2Cfg::PPL));
  DFS_PROPAGATE(make_act_tensor_map(&s->tmap[5], s->act[5], Dec3Cfg::WROWS, Dec3Cfg::WCOLS, Dec3Cfg::PPL));
  return DFS_OK;
}

// dec2's output_padding column (feature index 44 -> padded column 45) of d2 holds relu(folded bias) at every time step
__global__ void cae_fill_bias_column_kernel(ActBuf d2, int cols, int n_utts, const float* __restrict__ bias /*[64]*/) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_utts * 80 * 64;
  if (idx >= total) return;
  const int c = (int)(idx % 64);
  const int t = (int)((idx / 64) % 80);
  const long long n = idx / (64 * 80);
  const __half v = __float2half_rn(relu_nan(bias[c]));
  d2.ptr[(c >> 3) * d2.plane_elems() + ((n * cols + 45) * d2.RS + t + 1) * 8 + (c & 7)] = *reinterpret_cast<const uint16_t*>(&v);
}

int cae_tc_init_constants(CaeTcState* s, int max_utts, const float* dec2_bias_dev, cudaStream_t stream) {
  const long long total = (long long)max_utts * 80 * 64;
  cae_fill_bias_column_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, stream>>>(s->act[5], kCaeCols[5], max_utts, dec2_bias_dev);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

// ------------------------------------------------------------------------------------------
// final ConvTranspose2d(32,1,k2,s2) + zero row 320 + per-utterance MSE, one block per utterance
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float cae_in(const float* __restrict__ x, long long sn, long long st, long long sf, long long n, int t, int f,
                                        const float* __restrict__ mean, const float* __restrict__ sd) {
  float v = x[n * sn + t * st + f * sf];
  if (mean != nullptr) v = (v - mean[f]) / sd[f];
  return v;
}

// grid = (utterances, 20): block (n, s) handles the d3 rows to = 8s .. 8s+7 of all 90 columns, i.e. the 16 full input rows
// t = 16s .. 16s+15 (block 0 also the zero-padded row 320).
//   phase 1  thread -> (column fo, row to): eight consecutive threads read 128 contiguous bytes of each of the 4 channel
//            planes of d3; 4 outputs x 32 MACs; the 2x2 reconstruction patch goes to shared memory [16][180]
//   phase 2  the residual against the (normalised) input runs over that tile in the input's own storage order, so the
//            global reads are contiguous whichever of (t, f) is the fast axis (the first version read x through the
//            (to, fo) mapping: 4 useful bytes per 32-byte sector, and the kernel took as long as the 531 MFLOP layers).
// The block that arrives last adds the 20 partial sums in index order (deterministic score).
// The 4 x 32 weights travel in the kernel parameter space: every FFMA takes its weight as a constant-bank operand (no
// shared-memory loads, no 128 weight registers -- the first version needed 164 registers and ran one block per SM).
struct CaeFinalW {
  float w[128];   // [(a*2+b)*32 + ci]
  float bias;
};
__global__ void __launch_bounds__(256, 4) cae_final_tc_kernel(ActBuf d3, const float* __restrict__ x, long long sn, long long st, long long sf,
                                                               const float* __restrict__ mean, const float* __restrict__ sd,
                                                               const __grid_constant__ CaeFinalW fw,
                                                               float* __restrict__ mse_out, float* __restrict__ recon_out,
                                                               float* __restrict__ partial, unsigned int* __restrict__ done) {
  const long long n = blockIdx.x;
  const int split = blockIdx.y;
  __shared__ float rec[16][kF + 1];
  const long long plane_elems = d3.plane_elems();
  for (int pos = threadIdx.x; pos < 90 * 8; pos += blockDim.x) {
    const int fo = pos >> 3, tl = pos & 7, to = 8 * split + tl;
    const uint16_t* src = d3.ptr + ((n * 92 + fo + 1) * d3.RS + to + 1) * 8;
    float in[32];
#pragma unroll
    for (int pj = 0; pj < 4; ++pj) {
      const uint4 q = *reinterpret_cast<const uint4*>(src + pj * plane_elems);
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __half2 hh = *reinterpret_cast<const __half2*>(&u[e]);
        in[pj * 8 + 2 * e] = __low2float(hh);
        in[pj * 8 + 2 * e + 1] = __high2float(hh);
      }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        float r = fw.bias;
#pragma unroll
        for (int ci = 0; ci < 32; ++ci) r = fmaf(in[ci], fw.w[(a * 2 + b) * 32 + ci], r);
        rec[2 * tl + a][2 * fo + b] = r;
      }
  }
  __syncthreads();
  float acc = 0.0f;
  const int t0 = 16 * split;
  if (sf <= st) {   // feature axis is the fast one (contiguous [N,321,180]): consecutive threads -> consecutive features
    for (int i = threadIdx.x; i < 16 * kF; i += blockDim.x) {
      const int tt = i / kF, f = i - tt * kF;
      const float r = rec[tt][f];
      if (recon_out != nullptr) recon_out[n * kT * kF + (t0 + tt) * kF + f] = r;
      const float d = r - cae_in(x, sn, st, sf, n, t0 + tt, f, mean, sd);
      acc = fmaf(d, d, acc);
    }
  } else {          // time axis is the fast one (the reference's transposed view of [N,180,321] storage)
    for (int i = threadIdx.x; i < 16 * kF; i += blockDim.x) {
      const int f = i >> 4, tt = i & 15;
      const float r = rec[tt][f];
      if (recon_out != nullptr) recon_out[n * kT * kF + (t0 + tt) * kF + f] = r;
      const float d = r - cae_in(x, sn, st, sf, n, t0 + tt, f, mean, sd);
      acc = fmaf(d, d, acc);
    }
  }
  if (split == 0) {
    for (int f = threadIdx.x; f < kF; f += blockDim.x) {  // reconstruction row 320 is zero padding (model_cae.py:116-119)
      if (recon_out != nullptr) recon_out[n * kT * kF + 320 * kF + f] = 0.0f;
      const float d = cae_in(x, sn, st, sf, n, 320, f, mean, sd);
      acc = fmaf(d, d, acc);
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && mse_out != nullptr) {
    float s = 0.0f;
    for (int i = 0; i < 8; ++i) s += part[i];
    partial[n * kCaeFinalSplit + split] = s;
    __threadfence();
    if (atomicAdd(&done[n], 1u) == kCaeFinalSplit - 1) {   // last block of this utterance
      __threadfence();
      float tot = 0.0f;
      for (int i = 0; i < kCaeFinalSplit; ++i) tot += __ldcg(&partial[n * kCaeFinalSplit + i]);
      mse_out[n] = tot / (float)(kT * kF);
      done[n] = 0u;
    }
  }
}

// scores of the fused dec3 + final path: an utterance is exactly three 16-column units of the d2 layout (48 columns), so
// mse = (partial[3n] + partial[3n+1] + partial[3n+2] + sum_f x_norm[320][f]^2) / (321 * 180); the last term is the
// reconstruction's zero-padded row 320 (model_cae.py:116-119).  One warp per utterance, fixed summation order.
__global__ void __launch_bounds__(128) cae_mse_finish_kernel(const float* __restrict__ partial, const float* __restrict__ x, long long sn, long long st,
                                                              long long sf, const float* __restrict__ mean, const float* __restrict__ sd, int n_utts,
                                                              float* __restrict__ mse_out) {
  const int n = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= n_utts) return;
  float acc = 0.0f;
  for (int f = lane; f < kF; f += 32) {
    const float d = cae_in(x, sn, st, sf, n, 320, f, mean, sd);
    acc = fmaf(d, d, acc);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) mse_out[n] = (((partial[3 * n] + partial[3 * n + 1]) + partial[3 * n + 2]) + acc) / (float)(kT * kF);
}

// FT8 / FT8P -> [n][H][W][C] fp32 (H = time, W = feature): latent export and the per-layer debug dump
__global__ void ft8_unpack_kernel(ActBuf a, int cols, int parity_layout, int H, int W, int C, long long total, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const int wv = (int)((idx / C) % W);
  const int hv = (int)((idx / ((long long)C * W)) % H);
  const long long n = idx / ((long long)C * W * H);
  const long long gc = n * cols + wv + 1;
  long long off;
  if (parity_layout) off = (long long)((hv & 1) * (C / 8) + (c >> 3)) * a.plane_elems() + (gc * a.RS + (hv >> 1) + 1) * 8 + (c & 7);
  else off = (long long)(c >> 3) * a.plane_elems() + (gc * a.RS + hv + 1) * 8 + (c & 7);
  out[idx] = act_bits_to_float(a.ptr[off]);
}

static const int kCaeH[7] = {160, 80, 40, 20, 40, 80, 160};
static const int kCaeW[7] = {90, 45, 22, 11, 22, 45, 90};
static const int kCaeC[7] = {32, 64, 128, 256, 128, 64, 32};

int cae_tc_dump_layer(const CaeTcState* s, int layer, int n_utts, float* out_nhwc, cudaStream_t stream) {
  DFS_REQUIRE(layer >= 0 && layer < 7, DFS_ERR_INVALID, "CAE layer index %d out of range", layer);
  const long long total = (long long)n_utts * kCaeH[layer] * kCaeW[layer] * kCaeC[layer];
  if (total == 0) return DFS_OK;
  ft8_unpack_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, stream>>>(s->act[layer], kCaeCols[layer], layer == 0, kCaeH[layer], kCaeW[layer],
                                                                           kCaeC[layer], total, out_nhwc);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

// latent [n][256][20][11] (NCHW, the reference's return value) from e4
__global__ void cae_latent_nchw_kernel(ActBuf e4, int cols, long long total, float* __restrict__ latent) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int xw = (int)(idx % 11);
  const int y = (int)((idx / 11) % 20);
  const int c = (int)((idx / 220) % 256);
  const long long n = idx / (220 * 256);
  latent[idx] = act_bits_to_float(e4.ptr[(long long)(c >> 3) * e4.plane_elems() + ((n * cols + xw + 1) * e4.RS + y + 1) * 8 + (c & 7)]);
}

static ConvParams base_params(const CaeTcState* s, int li /*weights index 0..5*/, int in_layer, int out_layer, int n_utts, int feats, int rows_valid,
                              int out_feats) {
  ConvParams p{};
  p.wpack = s->w[li];
  for (int i = 0; i < 256; ++i) p.bias[i] = s->bias[li][i];
  p.n_units = num_col_tiles(n_utts, kCaeCols[in_layer]);
  p.n_utts = n_utts;
  p.cols = kCaeCols[in_layer];
  p.feats = feats;
  p.rows_valid = rows_valid;
  p.out = s->act[out_layer].ptr;
  p.out_ncols = s->act[out_layer].ncols;
  p.out_rs = s->act[out_layer].RS;
  p.out_cols = kCaeCols[out_layer];
  p.out_feats = out_feats;
  return p;
}

int launch_cae_tc(const CaeTcState* s, const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, const float* norm_mean, const float* norm_std,
                  float* mse_out, float* recon_out, float* latent_out, int stop_after_layer, int num_sms, cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  if (s->enc1_impl == 0)
    DFS_PROPAGATE(launch_cae_enc1_tc(x, sn, st, sf, n_utts, norm_mean, norm_std, s->xt1, s->w1pack, s->b1q, s->act[0], kCaeCols[0], num_sms, stream));
  else
    DFS_PROPAGATE(launch_conv1(x, sn, st, sf, n_utts, s->c1, norm_mean, norm_std, true, s->act[0], stream));
  if (stop_after_layer == 0) return DFS_OK;
  DFS_PROPAGATE(launch_conv_tc<Enc2Cfg>(s->tmap[0], base_params(s, 0, 0, 1, n_utts, 90, 80, 45), 1, num_sms, stream));
  if (stop_after_layer == 1) return DFS_OK;
  if (s->enc3_swap) {
    ConvParams p = base_params(s, 1, 1, 2, n_utts, 45, 80, 22);
    p.n_units = (int)(((long long)n_utts * kCaeCols[1] - 1 + Enc3SwapCfg::CT - 1) / Enc3SwapCfg::CT);
    DFS_PROPAGATE(launch_conv_tc<Enc3SwapCfg>(s->tmap_enc3_swap, p, 1, num_sms, stream));
  } else {
    DFS_PROPAGATE(launch_conv_tc<Enc3Cfg>(s->tmap[1], base_params(s, 1, 1, 2, n_utts, 45, 80, 22), 1, num_sms, stream));
  }
  if (stop_after_layer == 2) return DFS_OK;
  if (s->pair_mma)
    DFS_PROPAGATE(launch_conv_tc<Enc4PairCfg>(s->tmap[2], base_params(s, 2, 2, 3, n_utts, 22, 40, 11), 2, num_sms, stream));
  else
    DFS_PROPAGATE(launch_conv_tc<Enc4Cfg>(s->tmap[2], base_params(s, 2, 2, 3, n_utts, 22, 40, 11), 4, num_sms, stream));
  if (latent_out != nullptr) {
    const long long total = (long long)n_utts * 256 * 220;
    cae_latent_nchw_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, stream>>>(s->act[3], kCaeCols[3], total, latent_out);
    DFS_LAUNCH_CHECK();
  }
  if (stop_after_layer == 3) return DFS_OK;
  if (s->dec_wide) {
    ConvParams p1 = base_params(s, 3, 3, 4, n_utts, 11, 20, 22);
    p1.wpack = s->w_wide[0];
    DFS_PROPAGATE(launch_conv_tc<Dec1WideCfg>(s->tmap[3], p1, 2, num_sms, stream));
  } else {
    DFS_PROPAGATE(launch_conv_tc<Dec1Cfg>(s->tmap[3], base_params(s, 3, 3, 4, n_utts, 11, 20, 22), 4, num_sms, stream));
  }
  if (stop_after_layer == 4) return DFS_OK;
  if (s->dec_wide) {
    ConvParams p2 = base_params(s, 4, 4, 5, n_utts, 22, 40, 44);
    p2.wpack = s->w_wide[1];
    DFS_PROPAGATE(launch_conv_tc<Dec2WideCfg>(s->tmap[4], p2, 1, num_sms, stream));
  } else {
    DFS_PROPAGATE(launch_conv_tc<Dec2Cfg>(s->tmap[4], base_params(s, 4, 4, 5, n_utts, 22, 40, 44), 2, num_sms, stream));
  }
  if (stop_after_layer == 5) return DFS_OK;
  if (s->final_fused && stop_after_layer == 7 && mse_out != nullptr && recon_out == nullptr) {
    // scoring path: dec3's epilogue applies the final layer and accumulates the squared error; d3 never reaches HBM
    ConvParams p = base_params(s, 5, 5, 6, n_utts, 45, 80, 90);
    for (int q = 0; q < 4; ++q)
      for (int c = 0; c < 32; ++c) p.bias[32 + 4 * c + q] = s->w_final_host[q * 32 + c];   // channel-major: the pairs of an FFMA2 are adjacent
    p.bias[160] = s->final_bias;
    p.x = x;
    p.xsn = sn;
    p.xst = st;
    p.xsf = sf;
    p.norm_mean = norm_mean;
    p.norm_sd = norm_std;
    p.partial = s->mse_partial;
    p.x_vec4 = (sf == 1 && (st % 4) == 0 && (sn % 4) == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0) ? 1 : 0;
    DFS_PROPAGATE(launch_conv_tc<Dec3MseCfg>(s->tmap[5], p, 1, num_sms, stream));
    cae_mse_finish_kernel<<<(n_utts + 3) / 4, 128, 0, stream>>>(s->mse_partial, x, sn, st, sf, norm_mean, norm_std, n_utts, mse_out);
    DFS_LAUNCH_CHECK();
    return DFS_OK;
  }
  DFS_PROPAGATE(launch_conv_tc<Dec3Cfg>(s->tmap[5], base_params(s, 5, 5, 6, n_utts, 45, 80, 90), 1, num_sms, stream));
  if (stop_after_layer == 6) return DFS_OK;
  if (mse_out != nullptr || recon_out != nullptr) {
    CaeFinalW fw;
    memcpy(fw.w, s->w_final_host, sizeof(fw.w));
    fw.bias = s->final_bias;
    cae_final_tc_kernel<<<dim3(n_utts, kCaeFinalSplit), 256, 0, stream>>>(s->act[6], x, sn, st, sf, norm_mean, norm_std, fw, mse_out, recon_out,
                                                                         s->mse_partial, s->mse_done);
    DFS_LAUNCH_CHECK();
  }
  return DFS_OK;
}

}  // namespace dfs
