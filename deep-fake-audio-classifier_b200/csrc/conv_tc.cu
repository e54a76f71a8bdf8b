// conv_tc.cu -- the 2D-CNN's two tensor-core convolutions (instantiations of conv_tc.cuh) and the TMA
// tensor-map builder.  Replaces, on the scoring path, the reference's library calls
//   nn.Conv2d(32,64,3,p=1)+BatchNorm2d+ReLU+AvgPool2d((2,1))   /root/reference/src/model.py:21-24   (PAIR GEMM, EPI_PAIR_POOL)
//   nn.Conv2d(64,128,3,p=1)+BatchNorm2d+ReLU, x.mean(dim=2)    /root/reference/src/model.py:27-29,37 (3x3 GEMM, EPI_MEAN_T)
#include "conv_tc.cuh"

namespace dfs {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

// Tensor map over one FT8 / FT8P activation buffer: dim0 = (row, 8 channels) flattened and contiguous,
// dim1 = column, dim2 = plane; box = (wrows*8, wcols, box_planes).
int make_act_tensor_map(CUtensorMap* out, const ActBuf& a, int wrows, int wcols, int box_planes) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  DFS_REQUIRE(enc != nullptr, DFS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  DFS_REQUIRE(box_planes >= 1 && box_planes <= a.planes && wrows * 8 <= 256 && wcols <= 256, DFS_ERR_INVALID, "bad TMA box");
  cuuint64_t gdim[3] = {(cuuint64_t)a.RS * 8, (cuuint64_t)a.ncols, (cuuint64_t)a.planes};
  cuuint64_t gstr[2] = {(cuuint64_t)a.RS * 16, (cuuint64_t)a.ncols * a.RS * 16};
  cuuint32_t box[3] = {(cuuint32_t)wrows * 8, (cuuint32_t)wcols, (cuuint32_t)box_planes};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, a.ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DFS_REQUIRE(r == CUDA_SUCCESS, DFS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return DFS_OK;
}

// CNN2D conv2: 32 -> 64 channels on 160 x 180 as 80 time PAIRS per column, pooled to 80 rows.
using Conv2Cfg = ConvCfg<MODE_PAIR, 32, 64, 128, 80, 2, 3, 4, 1, EPI_PAIR_POOL>;
// (A CTA-pair variant of conv2 -- ConvCfg<..., CTA2 = 1>, M = 256, 64 weight rows per CTA -- was measured and removed: bit-identical, but
// 616 vs 985 TFLOP/s.  A cta_group::2 MMA costs the cycles of a single-CTA MMA of the same N (profiles/r01f_umma_bench_pairs.txt):
// each SM still streams its own 128 rows of A and consumes all N rows of B, so pairs only pay where shared memory forces N < 128.)
// CNN2D conv3: 64 -> 128 channels on 80 x 180, summed over time.
using Conv3Cfg = ConvCfg<MODE_3X3S, 64, 128, 256, 80, 1, 3, 2, 2, EPI_MEAN_T_SWAP>;   // weights as A, 256 positions as N
using Conv3PlainCfg = ConvCfg<MODE_3X3, 64, 128, 128, 80, 2, 2, 4, 1, EPI_MEAN_T>;   // positions as A (N = 128), kept for comparison

// "split" precision (option precision = 2; conv_tc.cuh SPLIT): value + residual planes and weight images, three MMAs per product.  Both
// run on CTA pairs so that the doubled weight images stay resident (98 KB / 147 KB per CTA).
using Conv2SplitCfg = ConvCfg<MODE_PAIR, 32, 64, 128, 80, 1, 4, 4, 1, EPI_PAIR_POOL, 1, 1>;
using Conv3SplitCfg = ConvCfg<MODE_3X3, 64, 128, 128, 80, 1, 3, 4, 1, EPI_MEAN_T, 1, 1>;

int make_cnn2d_split_tensor_maps(CUtensorMap* tmap_act1, CUtensorMap* tmap_act2, const ActBuf& act1, const ActBuf& act2) {
  DFS_PROPAGATE(make_act_tensor_map(tmap_act1, act1, Conv2SplitCfg::WROWS, Conv2SplitCfg::WCOLS, Conv2SplitCfg::PPL));
  DFS_PROPAGATE(make_act_tensor_map(tmap_act2, act2, Conv3SplitCfg::WROWS, Conv3SplitCfg::WCOLS, Conv3SplitCfg::PPL));
  return DFS_OK;
}

int make_cnn2d_tensor_maps(CUtensorMap* tmap_act1, CUtensorMap* tmap_act2, const ActBuf& act1, const ActBuf& act2) {
  DFS_PROPAGATE(make_act_tensor_map(tmap_act1, act1, Conv2Cfg::WROWS, Conv2Cfg::WCOLS, Conv2Cfg::PPL));
  DFS_PROPAGATE(make_act_tensor_map(tmap_act2, act2, Conv3Cfg::WROWS, Conv3Cfg::WCOLS, Conv3Cfg::PPL));
  return DFS_OK;
}

int launch_cnn2d_conv2_tc(const CUtensorMap& tmap_act1, const uint16_t* wpack, const float* bias_half, int n_utts, ActBuf act2,
                          int num_sms, cudaStream_t stream) {
  ConvParams p{};
  p.wpack = wpack;
  for (int i = 0; i < 64; ++i) p.bias[i] = bias_half[i];
  p.n_units = num_col_tiles(n_utts, kCols);
  p.n_utts = n_utts;
  p.cols = kCols;
  p.feats = kF;
  p.rows_valid = 80;
  p.out = act2.ptr;
  p.out_ncols = act2.ncols;
  p.out_rs = act2.RS;
  p.out_cols = kCols;
  p.out_feats = kF;
  return launch_conv_tc<Conv2Cfg>(tmap_act1, p, 1, num_sms, stream);
}

int launch_cnn2d_conv2_split(const CUtensorMap& tmap_act1, const uint16_t* wpack, const float* bias_half, float inv_scale, int n_utts, ActBuf act2,
                             int num_sms, cudaStream_t stream) {
  ConvParams p{};
  p.wpack = wpack;
  p.inv_scale = inv_scale;
  for (int i = 0; i < 64; ++i) p.bias[i] = bias_half[i];
  p.n_units = num_col_tiles(n_utts, kCols);
  p.n_utts = n_utts;
  p.cols = kCols;
  p.feats = kF;
  p.rows_valid = 80;
  p.out = act2.ptr;
  p.out_ncols = act2.ncols;
  p.out_rs = act2.RS;
  p.out_cols = kCols;
  p.out_feats = kF;
  return launch_conv_tc<Conv2SplitCfg>(tmap_act1, p, 1, num_sms, stream);
}

int launch_cnn2d_conv3_split(const CUtensorMap& tmap_act2, const uint16_t* wpack, const float* bias, float inv_scale, int n_utts, float* emb,
                             int num_sms, cudaStream_t stream) {
  ConvParams p{};
  p.wpack = wpack;
  p.inv_scale = inv_scale;
  for (int i = 0; i < 128; ++i) p.bias[i] = bias[i];
  p.n_units = num_col_tiles(n_utts, kCols);
  p.n_utts = n_utts;
  p.cols = kCols;
  p.feats = kF;
  p.rows_valid = 80;
  p.emb = emb;
  return launch_conv_tc<Conv3SplitCfg>(tmap_act2, p, 1, num_sms, stream);
}

int launch_cnn2d_conv3_tc(const CUtensorMap& tmap_act2, const uint16_t* wpack, const float* bias, int n_utts, float* emb,
                          int num_sms, cudaStream_t stream) {
  ConvParams p{};
  p.wpack = wpack;
  for (int i = 0; i < 128; ++i) p.bias[i] = bias[i];
  p.n_units = (int)(((long long)n_utts * kCols - 1 + Conv3Cfg::CT - 1) / Conv3Cfg::CT);
  p.n_utts = n_utts;
  p.cols = kCols;
  p.feats = kF;
  p.rows_valid = 80;
  p.emb = emb;
  return launch_conv_tc<Conv3Cfg>(tmap_act2, p, 1, num_sms, stream);
}

}  // namespace dfs
