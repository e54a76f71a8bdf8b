/* dfs_b200.h -- C ABI of the B200-native scoring engine (libdfs_b200.so).
 *
 * The reference (kingdomseed/Deep-Fake-Audio-Classifier) is pure Python and has no FFI; its
 * seams for this path are Python call signatures.  Each entry point below
 * names the reference interface it stands behind.  The Python host (dfs_b200/_native.py,
 * ctypes) is the only caller; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - plain C types only: pointers, sizes, ints.  No torch / C++ types cross this boundary.
 *  - every call returns an int status (DFS_OK = 0, < 0 = error); dfs_last_error() gives a
 *    thread-local message.  No exceptions, no callbacks.
 *  - "dev" pointers are CUDA device pointers owned by the caller (e.g. torch tensors);
 *    "host" pointers are host memory (pinned for the *_host entry points to overlap copies).
 *  - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = default stream).
 *    Calls that return results through host pointers synchronise that stream before returning.
 *  - a dfs_model owns its folded/re-packed weights and its activation workspace; it is bound
 *    to one device and is not thread-safe (one handle per device per thread, like the
 *    reference's single-threaded host loop, SURVEY.md §8b "Threading").
 *  - the metric calls (dfs_eer, dfs_eer_select, dfs_confusion, dfs_blend_f64, dfs_bce_with_logits) share one grow-only
 *    scratch workspace per device; the library serialises them per device with a mutex held for the whole call.  dfs_blend_f64
 *    and dfs_widen_f32_f64 return with their kernels still enqueued: issue the next metric call of that device on the SAME
 *    stream, or synchronise the stream first.
 *  - there is NO CPU fallback: without a CUDA device every compute call fails with
 *    DFS_ERR_CUDA.
 */
#ifndef DFS_B200_H
#define DFS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFS_OK 0
#define DFS_ERR_INVALID (-1)     /* bad argument (NULL, negative size, unsupported shape) */
#define DFS_ERR_CUDA (-2)        /* CUDA runtime / driver error, message has the detail   */
#define DFS_ERR_UNSUPPORTED (-3) /* architecture parameters the kernels are not built for */
#define DFS_ERR_NOMEM (-4)

#define DFS_T_FRAMES 321 /* time frames of an utterance map (README.md:76 of the reference) */
#define DFS_N_FEATS 180  /* LFCC + delta + delta-delta                                        */

typedef struct dfs_model dfs_model; /* opaque */

/* One Conv(+BatchNorm) block as the reference state_dict stores it (fp32, HOST pointers,
 * un-folded).  bn_* may all be NULL for a conv without BatchNorm (CAE decoder.9).
 * BN is folded in double precision at create time:  w' = w*g/sqrt(var+1e-5),
 * b' = (b-mean)*g/sqrt(var+1e-5)+beta.                                                    */
typedef struct {
  const float* weight; /* Conv2d (Co,Ci,3,3) | Conv1d (Co,Ci,3) | ConvTranspose2d (Ci,Co,2,2) */
  const float* bias;   /* (Co)                                                               */
  const float* bn_weight;
  const float* bn_bias;
  const float* bn_mean;
  const float* bn_var;
} dfs_conv_bn;

/* src/model.py:12-31  CNN2D(in_features=180, base_channels=32, num_classes=1) */
typedef struct {
  int in_features;    /* must be 180 */
  int base_channels;  /* must be 32  */
  dfs_conv_bn conv[3];    /* conv.{0,5,10} + BN conv.{1,6,11}        */
  const float* fc_weight; /* classifier.weight (1, 128*in_features)  */
  const float* fc_bias;   /* classifier.bias (1)                     */
} dfs_cnn2d_weights;

/* src/model_cnn1d.py:12-35  CNN1D(in_features=180, base_channels=32, num_classes=1) */
typedef struct {
  int in_features;
  int base_channels;
  dfs_conv_bn conv[3];    /* conv.{0,4,8} + BN conv.{1,5,9} */
  const float* fc_weight; /* (1, 128) */
  const float* fc_bias;
} dfs_cnn1d_weights;

/* src/model_cae.py:23-81  ConvAutoencoder(base_channels=32) (+ FeatureNormalizer stats,
 * src/dataset_cae.py:18-52; NULL = input is already normalised) */
typedef struct {
  int base_channels;
  dfs_conv_bn enc[4];     /* encoder.{0,4,8,12} + BN encoder.{1,5,9,13}                   */
  dfs_conv_bn dec[4];     /* decoder.{0,3,6,9} + BN decoder.{1,4,7}; dec[3] has no BN      */
  const float* norm_mean; /* (180) or NULL */
  const float* norm_std;  /* (180) or NULL */
} dfs_cae_weights;

/* src/dlqueen_model.py:132-173  DeepfakeDetector(in_ch=180, hidden=256): ConvEncoder (Conv1d k5 / k3 / k3 + BatchNorm1d +
 * GELU), masked mean+std StatsPool, head Linear(512,256) + GELU + Linear(256,1).  state_dict keys: enc.net.{0,4,8} convs,
 * enc.net.{1,5,9} BN, head.{0,3}.                                                                                   */
typedef struct {
  int in_ch;              /* must be 180 */
  int hidden;             /* must be 256 */
  dfs_conv_bn conv[3];    /* enc.net.{0,4,8} (256,in,5) / (256,256,3) / (256,256,3) + BN enc.net.{1,5,9} */
  const float* fc1_weight; /* head.0.weight (256, 512) */
  const float* fc1_bias;   /* head.0.bias (256)        */
  const float* fc2_weight; /* head.3.weight (1, 256)   */
  const float* fc2_bias;   /* head.3.bias (1)          */
} dfs_dlq_weights;

/* Strided view of the feature maps: element (i, t, f) is at x[i*stride_n + t*stride_t + f*stride_f]
 * (strides in ELEMENTS).  The reference hands its models a transposed, non-contiguous
 * (B,321,180) view of (B,180,321) storage (src/predict.py:103-105): stride_t = 1,
 * stride_f = 321 there; the benchmark layout is contiguous [N,321,180].                   */
typedef struct {
  const float* x;
  int64_t n;
  int64_t stride_n, stride_t, stride_f;
} dfs_features;

int dfs_version(void);
const char* dfs_last_error(void);
/* number of kernels this library has launched in the calling process (bench "gpu_launches") */
int64_t dfs_launch_count(void);

/* ---- model handles ------------------------------------------------------------------- */
/* `max_chunk` = utterances processed per internal pass (workspace is sized for it); 0 = default. */
int dfs_cnn2d_create(dfs_model** out, int device, const dfs_cnn2d_weights* w, int max_chunk);
int dfs_cnn1d_create(dfs_model** out, int device, const dfs_cnn1d_weights* w, int max_chunk);
int dfs_cae_create(dfs_model** out, int device, const dfs_cae_weights* w, int max_chunk);
int dfs_dlq_create(dfs_model** out, int device, const dfs_dlq_weights* w, int max_chunk);
int dfs_model_destroy(dfs_model* m);
/* options: "conv_impl" 0 = tcgen05 implicit GEMM (default), 1 = CUDA-core direct conv (debug
 * cross-check, same layouts); "profile" 0/1 = per-kernel event timing (dfs_model_profile);
 * "precision" (CNN2D, CNN1D, CAE) 0 = fp16 tensor-core operands with fp32 accumulation (default), 1 = the whole
 * network in fp32 on the CUDA cores (same arithmetic class as the reference's CPU path; for evaluations
 * where the rank order of scores a few 1e-6 apart matters, e.g. the EER of a small dev set), 2 (CNN2D) = "split": the
 * tensor-core path with every input sample, activation and weight carried as fp16 value + fp16 rounding residual and
 * three MMAs per product into the fp32 accumulator (fp32-class scores at about a third of the default rate).
 * Kernel-variant switches kept for on-device cross-checks (tests compare the variants; defaults are the product path):
 *   "conv1_impl"   (CNN2D, CAE) 0 = Toeplitz tcgen05 GEMM for the Cin = 1 layer, 1 = fp32 CUDA-core conv
 *   "conv12_fused" (CNN2D)      1 (default) = blocks 1 and 2 in ONE kernel (the layer-1 activations stay in shared memory), 0 = one
 *                               kernel per block (bit-identical to the pre-fusion path; act2 differs by one fp16 ulp on a few elements)
 *   "fused"        (CNN1D)      1 (default) = the three conv layers, the time mean and the classifier in ONE kernel (activations never
 *                               leave the SM; dense feature-contiguous input), 0 = one kernel per layer
 *   "l1_fused"     (CNN1D)      1 (default) = layer 1 converts the fp32 rows in flight, 0 = prep kernel + TMA
 *   "final_fused"  (CAE)        1 (default) = final ConvTranspose + squared error in dec3's epilogue, 0 = separate kernel over d3
 *   "dec_wide"     (CAE)        1 (default) = dec1 / dec2 as N = 256 GEMMs, 0 = N = 128 with twice the groups
 *   "enc3_swap"    (CAE)        1 (default) = enc3 with swapped operand roles (N = 256 positions), 0 = positions as M
 *   "pair_mma"     (CAE, StatsPool) 1 (default) = enc4 / layer 1 on CTA pairs (tcgen05 cta_group::2), 0 = single CTAs, N = 64 */
int dfs_model_set_option(dfs_model* m, const char* key, int64_t value);
int64_t dfs_model_workspace_bytes(const dfs_model* m);
/* With option "profile" = 1 every kernel launch of the scoring loop is bracketed by a CUDA event
 * pair on the launching stream.  This call synchronises those events and returns, per kernel id
 * (CNN2D: 0 conv1, 1 conv2, 2 conv3, 3 head), the summed device time in ms and the launch count
 * since the last reset.  Used by bench.py for the live roofline figure.                     */
int dfs_model_profile(dfs_model* m, double* ms_out, int64_t* launches_out, int n_ids, int reset);

/* Debug census of fp16 saturation.  Activations and the fp16 image of the features are stored with a saturating convert
 * (|v| > 65504 becomes +-65504 silently).  This scans the fp16 buffers the LAST pass (<= chunk utterances) of `m` left
 * behind -- feature image and every inter-layer activation -- and returns how many elements sit exactly at +-65504 and how
 * many are non-finite.  0 / 0 on real LFCC maps (range -61 ... +86, model_prediction_report.md:24-29); tests feed
 * heavy-tailed inputs and check it.  1D-CNN: needs option "fused" = 0 (the one-kernel path never stores its activations).
 * 2D-CNN: with the fused blocks 1 + 2 (default) the layer-1 activations stay in shared memory and are not part of the census
 * (set "conv12_fused" = 0 to include them); precision "split" counts the value planes, precision "fp32" returns 0 / 0.
 * Synchronises `stream`.                                                                                                   */
int dfs_model_saturation_count(dfs_model* m, int64_t* saturated_out, int64_t* nonfinite_out, void* stream);

/* ---- scoring (device-resident features) --------------------------------------------- */
/* CNN2D.forward + squeeze(-1) [+ torch.sigmoid]  (src/model.py:33-42, src/predict.py:106-108).
 * out_dev: [n] fp32 logits (apply_sigmoid=0) or scores.  embedding_dev: NULL or [n,23040] fp32
 * in the reference's flatten order c*180+f (src/model.py:38, return_embedding=True).       */
int dfs_cnn2d_score(dfs_model* m, const dfs_features* feats, float* out_dev, float* embedding_dev,
                    int apply_sigmoid, void* stream);
/* CNN1D.forward (src/model_cnn1d.py:37-46) */
int dfs_cnn1d_score(dfs_model* m, const dfs_features* feats, float* out_dev, int apply_sigmoid, void* stream);
/* get_cae_scores (src/predict_hybrid.py:66-78): per-utterance reconstruction MSE, the
 * reconstruction is never written to HBM.  apply_normalizer!=0 applies (x-mean)/std first
 * (src/predict_hybrid.py:45-49); the residual is taken against the normalised input.      */
int dfs_cae_score(dfs_model* m, const dfs_features* feats, int apply_normalizer, float* mse_dev, void* stream);
/* ConvAutoencoder.forward compat path (src/model_cae.py:83-125): materialises recon [n,321,180]
 * and latent [n,256,20,11] (either may be NULL).  Input must already be normalised.        */
int dfs_cae_forward(dfs_model* m, const dfs_features* feats, float* recon_dev, float* latent_dev, void* stream);

/* Debug: activations after CAE layer `layer` (0..6 = enc1..enc4, dec1..dec3) as [n][H][W][C] fp32 from the tensor-core
 * path (impl 0) or the CUDA-core cross-check path (impl 1); n <= the handle's chunk.  Tests only. */
int dfs_cae_debug_layer(dfs_model* m, const dfs_features* feats, int impl, int layer, int apply_normalizer, float* out_dev,
                        void* stream);

/* DeepfakeDetector.forward(x, lengths) (src/dlqueen_model.py:168-173; the reference stores x as (B, 180, T): pass the
 * (B, T, 180) view with stride_t = 1, stride_f = T).  T = 321; lengths_dev = NULL (all 321) or [n] int32 valid frame counts
 * for the masked pooling -- frames beyond an utterance's length must be zero in x, as pad_sequence leaves them
 * (src/dlqueen_model.py:98-103).  out_dev [n] logits (or sigmoid scores).                                            */
int dfs_dlq_score(dfs_model* m, const dfs_features* feats, const int32_t* lengths_dev, float* out_dev, int apply_sigmoid, void* stream);

/* ---- scoring (HOST features; copies are pipelined inside) --------------------------- */
/* The batch loop of src/predict.py:100-111 / src/predict_hybrid.py:52-78 behind one call:
 * feats->x and out_host are HOST pointers; H2D copies of chunk k+1 overlap the kernels of
 * chunk k; returns after the scores are in out_host.  `kind`: 0 cnn2d, 1 cnn1d, 2 cae-mse.
 * `flag` = apply_sigmoid (cnn) / apply_normalizer (cae).                                   */
int dfs_score_host(dfs_model* m, const dfs_features* feats, int flag, float* out_host, void* stream);
/* Same pipeline for a HOST slab of IEEE fp16 features, [n][321][180] (time_major = 0) or the reference's row shape
 * [n][180][321] (time_major = 1), dense: half the PCIe bytes.  The kernels quantise the features to fp16 before the first
 * GEMM anyway, so for the 2D-CNN and the 1D-CNN a slab holding the fp16 image of the fp32 features scores bit-identically;
 * the CAE additionally reads the input in its fp32 residual, so its MSE moves by the input rounding (~1e-4 relative). */
int dfs_score_host_f16(dfs_model* m, const uint16_t* x_host, int64_t n, int time_major, int flag, float* out_host, void* stream);

/* ---- scorer groups: ONE upload of the host table, every member model scores it --------- */
/* src/ensemble.py:105-122 and src/predict_hybrid.py:142-145 run one DataLoader pass over the feature table per model.
 * A group stages each slab of `stage_utts` utterances (0 = default 2,368; a single-model group keeps that model's pass
 * size) of the HOST table on the device once -- H2D of slab k+1 overlaps the kernels of slab k -- and every member
 * scores the staged slab before the buffer is re-used, so the PCIe bytes do not grow with the number of models.
 * The group borrows the model handles (they must outlive it and live on one device; not thread-safe, like the handles).
 * flags[i] = apply_sigmoid (2D-CNN, 1D-CNN, StatsPool detector) / apply_normalizer (CAE) of member i (NULL = all 1);
 * out_host[i] = HOST pointer to [n] fp32 scores of member i.  Returns after all scores are in host memory.            */
typedef struct dfs_group dfs_group; /* opaque */
int dfs_group_create(dfs_group** out, dfs_model* const* models, int n_models, int stage_utts);
int dfs_group_destroy(dfs_group* g);
int64_t dfs_group_stage_utts(const dfs_group* g);
int dfs_group_score_host(dfs_group* g, const dfs_features* feats_host, const int* flags, float* const* out_host, void* stream);
/* same for a dense HOST slab of IEEE fp16 features (layout as dfs_score_host_f16) */
int dfs_group_score_host_f16(dfs_group* g, const uint16_t* x_host, int64_t n, int time_major, const int* flags, float* const* out_host,
                             void* stream);

/* Page-locked host memory for the *_host entry points (cudaHostAlloc, portable; write_combined != 0 adds
 * cudaHostAllocWriteCombined: faster for the device to read over PCIe on some hosts, slow for the CPU to read back).
 * Stands where src/dataloaders.py:44-50 sets pin_memory=True on its DataLoaders.  Allocate after selecting the device.  */
int dfs_pinned_alloc(void** out_host, size_t bytes, int write_combined);
int dfs_pinned_free(void* p);

/* ---- ensemble blend (float64, like numpy) ------------------------------------------- */
/* out[i] = (sum_m weights[m] * (minmax_flags[m] ? normalise_01(scores[m])[i] : scores[m][i])) / divisor
 * evaluated left to right with separately rounded IEEE multiply / add / divide, like numpy:
 *   src/ensemble.py:121 np.mean(all_scores, axis=0)          -> weights 1, no min-max, divisor M
 *   src/predict_hybrid.py:81-85,149-151 alpha*a + (1-alpha)*b -> weights {alpha, 1-alpha}, min-max on, divisor 1
 * scores_host_array: m (<= 8) DEVICE pointers to [n] float64; out_dev [n] float64. */
int dfs_blend_f64(const double* const* scores_host_array, int m, const double* weights_host, const int* minmax_flags_host,
                  double divisor, int64_t n, double* out_dev, void* stream);
/* fp32 model scores -> float64 column (what .cpu().tolist() + np.array does, src/predict.py:111) */
int dfs_widen_f32_f64(const float* in_dev, int64_t n, double* out_dev, void* stream);

/* ---- EER (scripts/evaluation.py:7-39 == src/evaluation.py:12-48) -------------------- */
typedef struct {
  double eer;
  double threshold;
  int64_t eer_idx;    /* argmin index into the (n+1)-point FAR/FRR curves; -1 = single-class early-out */
  int64_t n_bonafide;
  int64_t n_spoof;
} dfs_eer_result;
/* scores_dev [n] fp32 or fp64 (key_bytes 4|8), labels_dev [n] uint8 in {0,1}.  Sort is a stable
 * LSD radix sort (ties keep original index order -- the kind="stable" contract of DESIGN.md);
 * perm_dev (NULL or [n] uint32) receives the permutation, sorted_dev (NULL or [n] same dtype as
 * the scores) the sorted scores.  Synchronises `stream`.                                    */
int dfs_eer(const void* scores_dev, int key_bytes, const uint8_t* labels_dev, int64_t n, dfs_eer_result* result_host,
            uint32_t* perm_dev, void* sorted_dev, void* stream);
/* Same (eer, threshold, eer_idx) without materialising the sort: FAR - FRR is strictly decreasing along the sorted
 * order, so the crossing is located by an MSD radix SELECT (one (digit, label) histogram per key byte, 5 B/score per
 * level) and only the tie group that holds it is put in stable order.  Bit-identical to dfs_eer on every input;
 * this is what calculate_eer(scores, labels) -> (eer, threshold) needs (scripts/evaluation.py:7-39).  Synchronises `stream`. */
int dfs_eer_select(const void* scores_dev, int key_bytes, const uint8_t* labels_dev, int64_t n, dfs_eer_result* result_host,
                   void* stream);
/* Library-wide switches for cross-checks (tests only; defaults are the product path):
 *   "eer_select_tma" 1 (default) = cp.async.bulk-fed histogram kernel, 0 = direct vector loads.
 *   "eer_sort_onesweep" (default 1): form of dfs_eer's radix passes.  0 = count / scan / scatter kernels over per-CTA
 *                       super-tiles (round 1); 1 = one scatter kernel per pass (tiles ticketed in input order, decoupled
 *                       look-back, next pass's histogram by shared-memory atomics in the same kernel), 512-thread tiles;
 *                       2 = the same on 256-thread tiles; 3 = histogram by a kernel of its own before each pass;
 *                       4 = histogram by ballots; 5 = the first pass reads the scores and labels itself.  All are stable
 *                       LSD sorts: identical permutation.
 *   "eer_sort_overlap" (default 1; fp32 scores, one-sweep forms 1 / 2 / 4): the pass over key byte 0 is launched before the
 *                       host has read back the label count and the key AND / OR (copied on a side stream), instead of
 *                       after that round trip; 0 = read back first.  Same result either way.                           */
int dfs_set_global_option(const char* key, int64_t value);
/* confusion_at_threshold (scripts/evaluation.py:42-56): out4_host = {tp, fp, tn, fn}. */
int dfs_confusion(const void* scores_dev, int key_bytes, const uint8_t* labels_dev, int64_t n, double threshold,
                  int64_t* out4_host, void* stream);

/* nn.BCEWithLogitsLoss() (mean) of a whole logit vector against float {0,1} labels -- the avg_loss of
 * src/evaluation.py::evaluate (:83-86,94), one fused reduction instead of a .item() per batch.
 * logits_dev / labels_dev [n] fp32 DEVICE pointers; *mean_host receives the loss.  Synchronises `stream`. */
int dfs_bce_with_logits(const float* logits_dev, const float* labels_dev, int64_t n, double* mean_host, void* stream);

/* ---- synthetic data (BASELINE.json north_star: "pinned synthetic feature tensors") --- */
/* Fill [n,321,180] fp32 with N(0, std^2) from a counter-based generator keyed by
 * (seed, first_utt + i, element) so every rank / GPU count sees the same global data set. */
int dfs_fill_features(float* out_dev, int64_t n, int64_t first_utt, uint64_t seed, float std, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFS_B200_H */
