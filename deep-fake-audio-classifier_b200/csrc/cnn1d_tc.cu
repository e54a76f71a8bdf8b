// cnn1d_tc.cu -- the 1D-CNN scorer on the tcgen05 template (conv_tc.cuh, MODE_3X1):
//   x.transpose(1,2) -> 3 x [Conv1d(k=3,p=1) + BatchNorm1d + ReLU] -> AdaptiveAvgPool1d(1) -> Linear(128,1)
//   /root/reference/src/model_cnn1d.py:37-46 (+ :14-35)
// Channels-last GEMMs over time: out[t,co] = b[co] + sum_k sum_f W[co,f,k] x[t+k-1,f].  One "feature column" per
// utterance (column index 1 + n), rows = time steps 1..321 padded to 328, so an MMA tile is 16 utterances x 8 time
// steps and the three taps are +0/+16/+32-byte shifts of the A descriptor.
//   prep    fp32 strided features -> FT8 fp16, 24 planes (180 features zero-padded to 192)
//   layer 1 K = 3 x 192, N = 64 (32 real output channels + 32 zero ones)      EPI_RELU
//   layer 2 K = 3 x 32 (the 32 real channels = planes 0..3 of the layer-1 buffer), N = 64   EPI_RELU
//   layer 3 K = 3 x 64, N = 128, time sum kept in registers over the 41 tiles   EPI_MEAN_T  -> [n][128] fp32
//   head    logits = fc_b + sum_c fc_w[c] * sum[c] / 321   (+ sigmoid)
// The arithmetic is ~31 MFLOP per utterance; the path is bound by the 231 KB/utterance fp32 input read.  On dense
// feature-contiguous input, prep + layer 1 are replaced by cnn1d_l1_fused.cu (fp32 -> fp16 conversion in flight).
#include "conv_tc.cuh"

namespace dfs {

constexpr int kC1dRows = 328;   // 321 time steps padded to a multiple of 8
constexpr int kC1dRS = 330;
using C1dL1 = ConvCfg<MODE_3X1, 192, 64, 64, kC1dRows, 1, 2, 4, 1, EPI_RELU>;
using C1dL2 = ConvCfg<MODE_3X1, 32, 64, 64, kC1dRows, 1, 3, 4, 1, EPI_RELU>;   // reads planes 0..3 of the layer-1 buffer only
using C1dL3 = ConvCfg<MODE_3X1, 64, 128, 128, kC1dRows, 1, 3, 4, 1, EPI_MEAN_T>;

void cnn1d_tc_geometry(int buf, int* planes, int* rs) {
  *planes = buf == 0 ? 24 : 8;
  *rs = kC1dRS;
}

int cnn1d_tc_make_maps(Cnn1dTcState* s) {
  DFS_PROPAGATE(make_act_tensor_map(&s->tmap[0], s->act[0], C1dL1::WROWS, C1dL1::WCOLS, C1dL1::PPL));
  DFS_PROPAGATE(make_act_tensor_map(&s->tmap[1], s->act[1], C1dL2::WROWS, C1dL2::WCOLS, C1dL2::PPL));
  DFS_PROPAGATE(make_act_tensor_map(&s->tmap[2], s->act[2], C1dL3::WROWS, C1dL3::WCOLS, C1dL3::PPL));
  return DFS_OK;
}

// item = (utterance, feature chunk of 8, time step); consecutive threads -> consecutive time steps (16-byte stores)
__global__ void __launch_bounds__(256) cnn1d_prep_kernel(const float* __restrict__ x, long long sn, long long st, long long sf, long long total,
                                                          ActBuf out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int t = (int)(idx % kT);
  const int c8 = (int)((idx / kT) % 24);
  const long long n = idx / ((long long)kT * 24);
  const float* src = x + n * sn + (long long)t * st;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int f = 8 * c8 + e;
    v[e] = f < kF ? src[(long long)f * sf] : 0.0f;   // pack_act2 saturates to +-65504
  }
  uint16_t* dst = out.ptr + (long long)c8 * out.plane_elems() + ((n + 1) * out.RS + t + 1) * 8;
  st_global_v4(dst, pack_act2(v[0], v[1]), pack_act2(v[2], v[3]), pack_act2(v[4], v[5]), pack_act2(v[6], v[7]));
}

// one warp per utterance
__global__ void __launch_bounds__(128) cnn1d_tc_head_kernel(const float* __restrict__ sums /*[n][128]*/, const float* __restrict__ fcw, float fcb,
                                                             int n_utts, int apply_sigmoid, float* __restrict__ out) {
  const int n = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= n_utts) return;
  float acc = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) acc = fmaf(sums[(long long)n * 128 + lane + 32 * k] / (float)kT, fcw[lane + 32 * k], acc);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const float z = acc + fcb;
    out[n] = apply_sigmoid ? 1.0f / (1.0f + expf(-z)) : z;
  }
}

static ConvParams c1d_params(const Cnn1dTcState* s, int layer, int n_utts) {
  ConvParams p{};
  p.wpack = s->w[layer];
  for (int i = 0; i < 128; ++i) p.bias[i] = s->bias[layer][i];
  p.n_units = (n_utts + kColTile - 1) / kColTile;   // columns 1 .. n_utts
  p.n_utts = n_utts;
  p.cols = 1;
  p.feats = 1;
  p.rows_valid = kT;
  if (layer < 2) {
    p.out = s->act[layer + 1].ptr;
    p.out_ncols = s->act[layer + 1].ncols;
    p.out_rs = s->act[layer + 1].RS;
  }
  p.out_cols = 1;
  p.out_feats = 1;
  p.emb = s->sums;
  return p;
}

int launch_cnn1d_tc(const Cnn1dTcState* s, const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, int apply_sigmoid, float* out, int num_sms,
                    cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  if (s->fused && cnn1d_l1_fused_supported(x, sn, st, sf))   // the whole network in one kernel: HBM traffic = the fp32 input read
    return launch_cnn1d_fused(x, sn, n_utts, s->w1_fused, s->w[1], s->w[2], s->bias[0], s->bias[1], s->bias[2], s->fcw_host, s->fcb, apply_sigmoid,
                              out, num_sms, stream);
  if (s->l1_fused && cnn1d_l1_fused_supported(x, sn, st, sf)) {
    DFS_PROPAGATE(launch_cnn1d_l1_fused(x, sn, n_utts, s->w[0], s->bias[0], s->act[1], num_sms, stream));
  } else {
    const long long total = (long long)n_utts * kT * 24;
    cnn1d_prep_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, stream>>>(x, sn, st, sf, total, s->act[0]);
    DFS_LAUNCH_CHECK();
    DFS_PROPAGATE(launch_conv_tc<C1dL1>(s->tmap[0], c1d_params(s, 0, n_utts), 1, num_sms, stream));
  }
  DFS_PROPAGATE(launch_conv_tc<C1dL2>(s->tmap[1], c1d_params(s, 1, n_utts), 1, num_sms, stream));
  DFS_PROPAGATE(launch_conv_tc<C1dL3>(s->tmap[2], c1d_params(s, 2, n_utts), 1, num_sms, stream));
  cnn1d_tc_head_kernel<<<(n_utts + 3) / 4, 128, 0, stream>>>(s->sums, s->fcw, s->fcb, n_utts, apply_sigmoid, out);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

// ==========================================================================================================
// StatsPool detector: /root/reference/src/dlqueen_model.py:115-173  (DeepfakeDetector, eval mode)
//   ConvEncoder  Conv1d(180,256,k5,p2)+BN+GELU -> Conv1d(256,256,k3,p1)+BN+GELU -> Conv1d(256,256,k3,p1)+BN+GELU
//   StatsPool    masked mean and std over time (two passes, var clamped at 1e-6)           -> (B, 512)
//   head         Linear(512,256) + GELU + Linear(256,1)
// The conv layers are the conv1d template with output-channel groups (the 256-wide weights of one layer are 390-480 KB):
// layer 1 (k = 5) as 4 groups of N = 64 (123 KB each), layers 2 and 3 as 2 groups of N = 128 (196 KB each), exact (erf)
// GELU in the epilogue.  Same input copy as
// the 1D-CNN (cnn1d_prep_kernel); the k = 5 halo rows beyond the stored pad row are the TMA's out-of-bounds zeros.
// Pooling + head are one block per utterance (256 threads = 256 channels / hidden units), fp32, fixed order.
using DlqL1 = ConvCfg<MODE_5X1, 192, 64, 64, kC1dRows, 1, 3, 4, 3, EPI_GELU>;
// the same layer on CTA pairs (option "pair_mma"): M = 256 = the two 16-utterance column tiles of a pair, 2 groups of N = 128 with 64
// weight rows (the same 123 KB group image) per CTA
using DlqL1Pair = ConvCfg<MODE_5X1, 192, 128, 128, kC1dRows, 1, 3, 4, 3, EPI_GELU, 1>;
static_assert(DlqL1Pair::WGT_B == DlqL1::WGT_B && DlqL1Pair::PPL == DlqL1::PPL && DlqL1Pair::WROWS == DlqL1::WROWS, "pair variant shares map and weights");
// layers 2 and 3: two groups of N = 128 (196 KB of resident weights; the activation window is cut into 8 pieces of 4 channel
// planes = 10 KB so that three stages still fit): an N = 128 MMA costs the same 88-100 cycles as an N = 64 one
using DlqL2 = ConvCfg<MODE_3X1, 256, 128, 128, kC1dRows, 1, 3, 4, 8, EPI_GELU>;

void dlq_geometry(int buf, int* planes, int* rs) {
  *planes = buf == 0 ? 24 : 32;
  *rs = kC1dRS;
}

int dlq_make_maps(DlqState* s) {
  DFS_PROPAGATE(make_act_tensor_map(&s->tmap[0], s->act0, DlqL1::WROWS, DlqL1::WCOLS, DlqL1::PPL));
  DFS_PROPAGATE(make_act_tensor_map(&s->tmap[1], s->actA, DlqL2::WROWS, DlqL2::WCOLS, DlqL2::PPL));
  DFS_PROPAGATE(make_act_tensor_map(&s->tmap[2], s->actB, DlqL2::WROWS, DlqL2::WCOLS, DlqL2::PPL));
  return DFS_OK;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// one block per utterance, thread c = encoder channel c / hidden unit c
__global__ void __launch_bounds__(256) dlq_stats_head_kernel(ActBuf h, const int32_t* __restrict__ lengths, const float* __restrict__ w1t /*[512][256]*/,
                                                              const float* __restrict__ b1, const float* __restrict__ w2, float b2, int apply_sigmoid,
                                                              float* __restrict__ out) {
  const int n = blockIdx.x, c = threadIdx.x;
  int len = lengths != nullptr ? lengths[n] : kT;
  len = len < 0 ? 0 : (len > kT ? kT : len);
  const uint16_t* col = h.ptr + (long long)(c >> 3) * h.plane_elems() + ((long long)(n + 1) * h.RS + 1) * 8 + (c & 7);
  const float denom = fmaxf((float)len, 1.0f);                      // mask.sum(dim=2).clamp(min=1.0)
  float s = 0.0f;
  for (int t = 0; t < len; ++t) s += act_bits_to_float(col[(long long)t * 8]);
  const float mean = s / denom;
  float v = 0.0f;
  for (int t = 0; t < len; ++t) {
    const float d = act_bits_to_float(col[(long long)t * 8]) - mean;
    v = fmaf(d, d, v);
  }
  const float sd = sqrtf(fmaxf(v / denom, 1e-6f));                  // torch.sqrt(var.clamp(min=1e-6))
  __shared__ float z[512];
  __shared__ float part[8];
  z[c] = mean;
  z[256 + c] = sd;
  __syncthreads();
  float a = b1[c];
#pragma unroll 8
  for (int k = 0; k < 512; ++k) a = fmaf(w1t[k * 256 + c], z[k], a);
  float contrib = gelu_erf(a) * w2[c];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
  if ((c & 31) == 0) part[c >> 5] = contrib;
  __syncthreads();
  if (c == 0) {
    float zl = b2;
    for (int k = 0; k < 8; ++k) zl += part[k];
    out[n] = apply_sigmoid ? 1.0f / (1.0f + expf(-zl)) : zl;
  }
}

static ConvParams dlq_params(const DlqState* s, int layer, int n_utts, const ActBuf& outbuf) {
  ConvParams p{};
  p.wpack = s->w[layer];
  for (int i = 0; i < 256; ++i) p.bias[i] = s->bias[layer][i];
  p.n_units = (n_utts + kColTile - 1) / kColTile;
  p.n_utts = n_utts;
  p.cols = 1;
  p.feats = 1;
  p.rows_valid = kT;
  p.out = outbuf.ptr;
  p.out_ncols = outbuf.ncols;
  p.out_rs = outbuf.RS;
  p.out_cols = 1;
  p.out_feats = 1;
  return p;
}

int launch_dlq(const DlqState* s, const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, const int32_t* lengths_dev, int apply_sigmoid,
               float* out, int num_sms, cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  const long long total = (long long)n_utts * kT * 24;
  cnn1d_prep_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, stream>>>(x, sn, st, sf, total, s->act0);
  DFS_LAUNCH_CHECK();
  if (s->pair_mma)
    DFS_PROPAGATE(launch_conv_tc<DlqL1Pair>(s->tmap[0], dlq_params(s, 0, n_utts, s->actA), 2, num_sms, stream));
  else
    DFS_PROPAGATE(launch_conv_tc<DlqL1>(s->tmap[0], dlq_params(s, 0, n_utts, s->actA), 4, num_sms, stream));
  DFS_PROPAGATE(launch_conv_tc<DlqL2>(s->tmap[1], dlq_params(s, 1, n_utts, s->actB), 2, num_sms, stream));
  DFS_PROPAGATE(launch_conv_tc<DlqL2>(s->tmap[2], dlq_params(s, 2, n_utts, s->actA), 2, num_sms, stream));
  dlq_stats_head_kernel<<<n_utts, 256, 0, stream>>>(s->actA, lengths_dev, s->fc1_wt, s->fc1_b, s->fc2_w, s->fc2_b, apply_sigmoid, out);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

}  // namespace dfs
