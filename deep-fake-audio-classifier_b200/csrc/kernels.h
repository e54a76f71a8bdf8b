// kernels.h -- host-callable launchers shared between the .cu translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dfs_b200.h"
#include "layout.cuh"

namespace dfs {

// ---- conv_tc.cu (tcgen05 implicit GEMM) ----
int make_act_tensor_map(CUtensorMap* out, const ActBuf& a, int wrows, int wcols, int box_planes);
int make_cnn2d_tensor_maps(CUtensorMap* tmap_act1, CUtensorMap* tmap_act2, const ActBuf& act1, const ActBuf& act2);
int launch_cnn2d_conv2_tc(const CUtensorMap& tmap_act1, const uint16_t* wpack, const float* bias, int n_utts, ActBuf act2,
                          int num_sms, cudaStream_t stream);
int launch_cnn2d_conv3_tc(const CUtensorMap& tmap_act2, const uint16_t* wpack, const float* bias, int n_utts, float* emb,
                          int num_sms, cudaStream_t stream);
// "split" precision (option precision = 2): value + residual planes / weight images, CTA pairs
int make_cnn2d_split_tensor_maps(CUtensorMap* tmap_act1, CUtensorMap* tmap_act2, const ActBuf& act1, const ActBuf& act2);
int launch_cnn2d_conv2_split(const CUtensorMap& tmap_act1, const uint16_t* wpack, const float* bias, float inv_scale, int n_utts, ActBuf act2,
                             int num_sms, cudaStream_t stream);
int launch_cnn2d_conv3_split(const CUtensorMap& tmap_act2, const uint16_t* wpack, const float* bias, float inv_scale, int n_utts, float* emb,
                             int num_sms, cudaStream_t stream);

// ---- conv1_tc.cu (CNN2D block 1 as a Toeplitz-in-time tcgen05 GEMM) ----
int64_t conv1_xt_rows(int64_t n_utts);   // 16-byte rows of the fp16 time-major feature copy for n utterances
int launch_conv1_prep(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, uint16_t* xt, cudaStream_t stream);
// conv12_fused.cu: conv1 + conv2 in one kernel (act1 stays in shared memory); weight images: see Conv12Params
int launch_cnn2d_conv12_fused(const uint16_t* xt, const uint16_t* w1pack, const float* b1_half, const uint16_t* w2pack, const float* b2_half, int n_utts,
                              ActBuf act2, int num_sms, cudaStream_t stream);
int launch_conv1_tc(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, uint16_t* xt, const uint16_t* wpack, const float* bias_half,
                    ActBuf out, int num_sms, cudaStream_t stream);
// split precision: xt_lo = the fp16 rounding residuals of xt (same geometry), out = 16 planes (8 value planes, then 8 residual planes)
int launch_conv1_tc_split(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, uint16_t* xt, uint16_t* xt_lo, const uint16_t* wpack,
                          float inv_scale, ActBuf out, int num_sms, cudaStream_t stream);

// ---- cnn2d.cu (CUDA-core stages of the 2D-CNN) ----
struct Conv1Weights {
  float w[32 * 9];  // folded, [co][kh][kw]
  float b[32];
};
// conv1 + BN + ReLU + AvgPool (2,1) [or (2,2) for the CAE encoder] -> FT8 fp16.
// x element (i,t,f) at x[i*sn + t*st + f*sf]; optional per-feature normaliser (mean, 1/std) applied first.
int launch_conv1(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, const Conv1Weights& w, const float* norm_mean,
                 const float* norm_std, bool pool_f, ActBuf out, cudaStream_t stream);
// debug cross-check of the tensor-core path: same inputs / packed weights / outputs, CUDA cores only
int launch_cnn2d_conv2_simt(ActBuf act1, const uint16_t* wpack, const float* bias_dev, int n_utts, ActBuf act2, cudaStream_t stream);
int launch_cnn2d_conv3_simt(ActBuf act2, const uint16_t* wpack, const float* bias_dev, int n_utts, float* emb, cudaStream_t stream);
// logits[n] = fc_b + sum_{f,c} emb[n][f][c] * wfc[f][c]  (wfc already carries 1/T); optional sigmoid
int launch_cnn2d_head(const float* emb, const float* wfc, float fcb, int n_utts, int apply_sigmoid, float* out, cudaStream_t stream,
                      bool acc64 = false);
// embedding[n][c*180+f] = emb[n][f][c] / 80   (src/model.py:37-38 flatten order)
int launch_cnn2d_embedding_export(const float* emb, int n_utts, float* embedding, cudaStream_t stream);

// ---- cnn2d_fp32.cu (option "precision" = 1: the whole 2D-CNN in fp32 on the CUDA cores) ----
size_t cnn2d_fp32_work_floats(int n_utts);
int launch_cnn2d_fp32(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, const float* const w[3], const float* const b[3],
                      const float* wfc, float fcb, int apply_sigmoid, float* work, float* emb, float* out, cudaStream_t stream);

// ---- cae_tc.cu (convolutional autoencoder on the tcgen05 template) ----
struct CaeTcState {
  Conv1Weights c1;        // folded encoder block 1 (fp32, CUDA cores)
  ActBuf act[7];          // e1 e2 e3 e4(latent) d1 d2 d3
  CUtensorMap tmap[6];    // inputs of enc2 enc3 enc4 dec1 dec2 dec3
  const uint16_t* w[6];   // packed fp16 weights of those six layers (device)
  float bias[6][256];     // folded biases per output channel, pre-scaled like the weights
  const float* w_final;   // final ConvTranspose2d(32,1): [(a*2+b)*32 + ci] fp32 (device)
  float final_bias;
  float w_final_host[128];// same weights on the host: passed to the fused final kernel as a kernel parameter (constant bank)
  uint16_t* xt1;          // enc1 input image xT2 (cae_enc1_tc.cu), zero padded
  const uint16_t* w1pack; // enc1 Toeplitz weights [kw][K chunk][n 256][8] fp16, 0.25 folded
  float b1q[32];          // 0.25 * folded enc1 bias
  int enc1_impl;          // 0 = tensor-core Toeplitz GEMM, 1 = fp32 CUDA-core conv1_kernel<POOLF> (cross-check)
  const uint16_t* w_wide[2]; // dec1 / dec2 weights for the N = 256 variants (option "dec_wide")
  CUtensorMap tmap_enc3_swap; // e2 with the 34-column x 10-row box of the swapped-role enc3
  int enc3_swap;          // 1 = enc3 with swapped operand roles (N = 256 positions), 0 = positions as M (N = 128)
  int pair_mma;           // 1 (default) = enc4 on CTA pairs (cluster of 2, tcgen05 cta_group::2: M = 256, 64 of the 128 weight rows per CTA)
  int dec_wide;           // 1 = dec1 / dec2 as N = 256 GEMMs (half the TMA re-reads of the input, one CTA per SM), 0 = N = 128
  int final_fused;        // 1 (default) = final layer + squared error in dec3's epilogue, 0 = separate cae_final_tc_kernel over d3
  float* mse_partial;     // [chunk][kCaeFinalSplit] partial squared-error sums of the fused final layer
  unsigned int* mse_done; // [chunk] arrival counters (self-resetting)
};
constexpr int kCaeFinalSplit = 20;  // blocks per utterance in the fused final ConvT + MSE kernel: 8 of the 160 d3 rows each
// ---- cae_enc1_tc.cu ----
int64_t cae_enc1_xt_rows(int64_t n_utts);
int launch_cae_enc1_tc(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, const float* norm_mean, const float* norm_std, uint16_t* xt,
                       const uint16_t* wpack, const float* bias_quarter, ActBuf out, int out_cols, int num_sms, cudaStream_t stream);
void cae_tc_geometry(int layer, int* planes, int* cols, int* rs);
int cae_tc_make_maps(CaeTcState* s);
int cae_tc_init_constants(CaeTcState* s, int max_utts, const float* dec2_bias_dev, cudaStream_t stream);
int cae_tc_dump_layer(const CaeTcState* s, int layer, int n_utts, float* out_nhwc, cudaStream_t stream);
// stop_after_layer: 0..6 = stop after e1..d3 (debug), 7 = run the fused final + MSE
int launch_cae_tc(const CaeTcState* s, const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, const float* norm_mean, const float* norm_std,
                  float* mse_out, float* recon_out, float* latent_out, int stop_after_layer, int num_sms, cudaStream_t stream);

// ---- cnn1d_tc.cu (1D-CNN on the tcgen05 template) ----
struct Cnn1dTcState {
  ActBuf act[3];          // fp16 input copy (24 planes), layer-1 output, layer-2 output (8 planes each)
  CUtensorMap tmap[3];
  const uint16_t* w[3];   // packed fp16 weights [tap][ci/8][co][8] (channel counts zero-padded to 192/64, 32/64, 64/128)
  float bias[3][128];
  float* sums;            // [n][128] time sums of the last layer
  const float* fcw;       // classifier weight (128) on the device
  float fcb;
  int l1_fused;           // 1 (default) = layer 1 converts the fp32 rows in flight (cnn1d_l1_fused.cu) when the layout allows
  int fused;              // 1 (default) = the whole network in ONE kernel (cnn1d_fused.cu) when the layout allows; 0 = one kernel per layer
  const uint16_t* w1_fused;  // layer-1 weights with 32 rows per K chunk, [tap][24][32][8] (the template's image pads them to 64)
  float fcw_host[128];    // classifier weight on the host: rides in the fused kernel's parameter space
};
// ---- StatsPool detector (dlqueen_model.py) on the conv1d template, cnn1d_tc.cu ----
struct DlqState {
  ActBuf act0;            // fp16 input copy, 24 planes (180 features zero-padded to 192)
  ActBuf actA, actB;      // 32-plane (256-channel) ping-pong activation buffers
  CUtensorMap tmap[3];    // inputs of the three conv layers: act0, actA, actB
  const uint16_t* w[3];   // packed fp16 weights [group][tap][ci/8][gco][8], gco = 64 (layer 1) / 128 (layers 2, 3)
  float bias[3][256];
  const float* fc1_wt;    // head.0.weight transposed to [512][256] (device)
  const float* fc1_b;     // [256]
  const float* fc2_w;     // head.3.weight [256]
  float fc2_b;
  int pair_mma;           // 1 (default) = layer 1 on CTA pairs (tcgen05 cta_group::2): 2 groups of N = 128 instead of 4 of N = 64
};
void dlq_geometry(int buf, int* planes, int* rs);
int dlq_make_maps(DlqState* s);
int launch_dlq(const DlqState* s, const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, const int32_t* lengths_dev, int apply_sigmoid,
               float* out, int num_sms, cudaStream_t stream);

// ---- cnn1d_fused.cu: the three conv layers, the time mean and the classifier in one kernel (dense fp32 input, as cnn1d_l1_fused) ----
int launch_cnn1d_fused(const float* x, int64_t sn, int n_utts, const uint16_t* w1, const uint16_t* w2, const uint16_t* w3, const float* b1,
                       const float* b2, const float* b3, const float* fcw_host, float fcb, int apply_sigmoid, float* out, int num_sms,
                       cudaStream_t stream);
// ---- cnn1d_l1_fused.cu ----
bool cnn1d_l1_fused_supported(const float* x, int64_t sn, int64_t st, int64_t sf);
int launch_cnn1d_l1_fused(const float* x, int64_t sn, int n_utts, const uint16_t* wpack, const float* bias, ActBuf out, int num_sms, cudaStream_t stream);
void cnn1d_tc_geometry(int buf, int* planes, int* rs);
int cnn1d_tc_make_maps(Cnn1dTcState* s);
int launch_cnn1d_tc(const Cnn1dTcState* s, const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, int apply_sigmoid, float* out, int num_sms,
                    cudaStream_t stream);

// ---- simt_models.cu (CUDA-core CNN1D and CAE) ----
struct SimtConv {        // BN-folded fp32 weights on device, re-packed [tap][ci][co] (co fastest)
  float* w = nullptr;
  float* b = nullptr;    // [co] folded bias
  int ci = 0, co = 0;
};
int launch_cnn1d_simt(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, const SimtConv* conv3, const float* fcw,
                      float fcb, int apply_sigmoid, float* work, float* out, cudaStream_t stream);
size_t cnn1d_simt_work_floats(int n_utts);
int launch_cae_simt(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, const SimtConv* enc4, const SimtConv* dec4,
                    float final_bias, const float* norm_mean, const float* norm_std, float* work, float* mse_out, float* recon_out,
                    float* latent_out, cudaStream_t stream);
size_t cae_simt_work_floats(int n_utts);
const float* cae_simt_layer_ptr(const float* work, int n_utts, int layer, size_t* floats_per_utt);

// ---- eer.cu ----
int eer_device(const void* scores, int key_bytes, const uint8_t* labels, int64_t n, dfs_eer_result* result_host, uint32_t* perm,
               void* sorted, cudaStream_t stream);
int eer_select_device(const void* scores, int key_bytes, const uint8_t* labels, int64_t n, dfs_eer_result* result_host, cudaStream_t stream);
int confusion_device(const void* scores, int key_bytes, const uint8_t* labels, int64_t n, double thr, int64_t* out4_host,
                     cudaStream_t stream);
int blend_device(const double* const* scores, int m, const double* weights, const int* minmax, double divisor, int64_t n, double* out,
                 cudaStream_t stream);
int bce_with_logits_device(const float* logits, const float* labels, int64_t n, double* mean_host, cudaStream_t stream);
int widen_device(const float* in, int64_t n, double* out, cudaStream_t stream);

// ---- synth.cu ----
int fill_features_device(float* out, int64_t n, int64_t first_utt, uint64_t seed, float std, cudaStream_t stream);

// ---- probe.cu (lib/libdfs_b200_probes.so only) ----
int probe_umma(const uint16_t* a, const uint16_t* b, int rows_a, int n, int k, int row_shift, int group_rows, float* out,
               cudaStream_t stream);
int probe_umma_bench(int n, int nmma, int iters, int n_acc, const uint32_t* a_off, const uint32_t* b_off, uint32_t a_lbo, uint32_t a_sbo,
                     uint32_t b_lbo, uint32_t b_sbo, uint32_t layout, uint32_t use_base_offset, long long* cycles_host, cudaStream_t stream);
int probe_tmem_ld_bench(int shape, int nwarps, int blocks, int iters, int lds_per_wait, long long* cycles_host, long long* bytes_per_block_host,
                        cudaStream_t stream);
int probe_tma_window(const uint16_t* act, int planes, int RS, int64_t ncols, int wrows, int row0, int col0, uint16_t* out,
                     cudaStream_t stream);

}  // namespace dfs
