// conv_tc.cu -- 3x3 convolution (+folded BN +ReLU +pool / +time-mean) as an implicit GEMM on the
// 5th-gen tensor cores: tcgen05.mma with fp32 accumulators in TMEM, operands staged by TMA.
//
// Replaces, on the scoring path, the reference's library calls
//   nn.Conv2d(32,64,3,p=1)+BatchNorm2d+ReLU+AvgPool2d((2,1))   /root/reference/src/model.py:21-24   (EPI_POOL_T)
//   nn.Conv2d(64,128,3,p=1)+BatchNorm2d+ReLU, x.mean(dim=2)    /root/reference/src/model.py:27-29,37 (EPI_MEAN_T)
//
// GEMM view (see layout.cuh for the activation layout):
//   one MMA tile  = 128 output positions = 16 feature columns x 8 time steps  (M = 128)
//   N             = COUT (64 / 128), K = 9 taps x CIN, issued as 9*CIN/16 tcgen05.mma of K = 16
//   A (activations): SWIZZLE_NONE K-major smem descriptor straight into the TMA-loaded window;
//                    tap (kh,kw) = +((kw*WROWS + kh) * 16) bytes on the start address
//   B (weights)    : BN-folded fp16, resident in shared memory for the whole kernel
//   D              : TMEM, NACC accumulators of COUT columns, so the epilogue of tile i overlaps
//                    the MMAs of tile i+1.
// Warp roles (352 threads): warps 0..7 = epilogue (TMEM lane quarter = warp%4, column half = warp/4),
// warp 8 = TMA producer, warp 9 = MMA issuer (one lane; highest warp id on its scheduler so the
// hi-warp-id-first arbiter never starves it behind spinning epilogue warps), warp 10 = TMEM allocator.
// Work unit = one column tile (16 feature columns, all T time steps); units are dealt round-robin
// to a persistent grid of one CTA per SM.
#include "common.cuh"
#include "kernels.h"
#include "layout.cuh"

namespace dfs {

enum { EPI_POOL_T = 0, EPI_MEAN_T = 1 };

template <int CIN_, int COUT_, int T_, int MT_, int NSTAGE_, int NACC_, int EPI_>
struct ConvCfg {
  static constexpr int CIN = CIN_, COUT = COUT_, T = T_, MT = MT_, NSTAGE = NSTAGE_, NACC = NACC_, EPI = EPI_;
  static constexpr int KCH = CIN / 8;                  // 16-byte K chunks = activation planes
  static constexpr int WROWS = 8 * MT + 2;             // window rows (time) incl. halo
  static constexpr int WCOLS = kColTile + 2;           // window columns (feature) incl. halo
  static constexpr int PLANE_B = WCOLS * WROWS * 16;   // bytes of one plane of the window
  static constexpr int WIN_B = KCH * PLANE_B;          // TMA transaction bytes per window
  static constexpr int WIN_B_AL = (WIN_B + 1023) & ~1023;
  static constexpr int WGT_B = 9 * CIN * COUT * 2;
  static constexpr int WGT_B_AL = (WGT_B + 1023) & ~1023;
  static constexpr int ST = T / (8 * MT);              // windows (super-tiles) per unit
  static constexpr int TILES = T / 8;                  // MMA tiles per unit
  static constexpr int TMEM_COLS = NACC * COUT;
  static constexpr int BAR_B = 256;
  static constexpr int SMEM_B = WGT_B_AL + NSTAGE * WIN_B_AL + BAR_B;
  static constexpr int THREADS = 352;
  // two CTAs per SM when shared memory and TMEM allow: the second CTA's epilogue / issue latencies
  // overlap the first one's MMAs (conv2: 101 KB smem, 256 TMEM columns each)
  static constexpr int OCC = (SMEM_B <= 113 * 1024 && TMEM_COLS <= 256) ? 2 : 1;
  static_assert(T % (8 * MT) == 0, "T must be a multiple of the super-tile height");
  static_assert(NACC % MT == 0, "the accumulators of one window must be consecutive");
  static_assert(WROWS * 8 <= 256, "TMA box inner dimension limit");
  static_assert(TMEM_COLS == 32 || TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns");
  static_assert(COUT % 64 == 0 && COUT <= 256 && CIN % 16 == 0, "shape");
  static_assert(SMEM_B <= 227 * 1024, "shared memory budget");
};

struct ConvParams {
  const uint16_t* wpack;  // [9][CIN/8][COUT][8] fp16, BN folded
  float bias[128];        // folded bias per output channel
  int n_units;            // column tiles
  int n_utts;
  int cols;               // padded feature columns per utterance (F + 2)
  int feats;              // F
  // EPI_POOL_T: pooled fp16 activations in FT8 layout with RS = T/2 + 2
  uint16_t* out;
  long long out_ncols;
  // EPI_MEAN_T: per-utterance time SUMS, [n][F][COUT] fp32 (the head applies 1/T)
  float* emb;
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::OCC)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ConvParams p) {
  constexpr int CIN = Cfg::CIN, COUT = Cfg::COUT, MT = Cfg::MT, NSTAGE = Cfg::NSTAGE, NACC = Cfg::NACC;
  constexpr int KCH = Cfg::KCH, WROWS = Cfg::WROWS, PLANE_B = Cfg::PLANE_B;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* wsm = smem;
  uint8_t* win0 = smem + Cfg::WGT_B_AL;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::WGT_B_AL + NSTAGE * Cfg::WIN_B_AL);
  uint64_t* full = bars;                    // [NSTAGE]  TMA -> MMA
  uint64_t* empty = bars + NSTAGE;          // [NSTAGE]  MMA -> TMA
  uint64_t* tfull = bars + 2 * NSTAGE;      // [NACC]    MMA -> epilogue
  uint64_t* tempty = tfull + NACC;          // [NACC]    epilogue -> MMA
  uint64_t* wbar = tempty + NACC;           // weights resident
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap);
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == 10) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(wbar, Cfg::WGT_B);
      constexpr int PIECE = 16384;
      for (int off = 0; off < Cfg::WGT_B; off += PIECE) {
        const int bytes = (Cfg::WGT_B - off) < PIECE ? (Cfg::WGT_B - off) : PIECE;
        bulk_g2s(wsm + off, reinterpret_cast<const uint8_t*>(p.wpack) + off, bytes, wbar);
      }
      uint32_t ws = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        for (int st = 0; st < Cfg::ST; ++st, ++ws) {
          const int stage = ws % NSTAGE;
          mbar_wait(&empty[stage], ((ws / NSTAGE) & 1) ^ 1, 1);
          mbar_arrive_expect_tx(&full[stage], Cfg::WIN_B);
          tma_load_3d(win0 + stage * Cfg::WIN_B_AL, &tmap, st * MT * 64, u * kColTile, 0, &full[stage]);
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(128, COUT);
      const uint32_t wsm_a = smem_u32(wsm);
      // descriptor = (low word: start address >> 4 | LBO >> 4 << 16, high word: SBO >> 4 | version); taps and
      // K steps only move the start address, i.e. add a compile-time constant to the low word
      const uint64_t b_desc0 = umma_smem_desc(wsm_a, COUT * 16, 128);
      const uint32_t b_lo0 = (uint32_t)b_desc0, b_hi = (uint32_t)(b_desc0 >> 32);
      const uint64_t a_desc0 = umma_smem_desc(smem_u32(win0), PLANE_B, WROWS * 16);
      const uint32_t a_lo0 = (uint32_t)a_desc0, a_hi = (uint32_t)(a_desc0 >> 32);
      mbar_wait(wbar, 0, 2);
      uint32_t ws = 0, it = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        for (int st = 0; st < Cfg::ST; ++st, ++ws) {
          const int stage = ws % NSTAGE;
          mbar_wait(&full[stage], (ws / NSTAGE) & 1, 3);
          tc_fence_after();
          const uint32_t a_lo_stage = a_lo0 + (uint32_t)(stage * (Cfg::WIN_B_AL >> 4));
          // The MT tiles of a window are issued INTERLEAVED (tile index innermost): consecutive MMAs then
          // accumulate into different TMEM tiles, so the ~90-cycle accumulate latency of one MMA (measured,
          // tools/umma_bench.py) is hidden behind the next one instead of serialising the K loop.
          const int acc0 = it % NACC;  // NACC % MT == 0: the MT accumulators of a window are consecutive
#pragma unroll
          for (int m = 0; m < MT; ++m) mbar_wait(&tempty[acc0 + m], (((it + m) / NACC) & 1) ^ 1, 4);
          tc_fence_after();
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int kh = tap / 3, kw = tap % 3;
#pragma unroll
            for (int kk = 0; kk < CIN / 16; ++kk) {
              const uint32_t a_off = (uint32_t)(((2 * kk) * PLANE_B + (kw * WROWS + kh) * 16) >> 4);
              const uint32_t b_off = (uint32_t)((((tap * KCH + 2 * kk) * COUT) * 16) >> 4);
#pragma unroll
              for (int m = 0; m < MT; ++m)  // tile m = rows 8m.. of the window: +8 rows of 16 B
                umma_f16_lohi(tmem_base + (acc0 + m) * COUT, a_lo_stage + (uint32_t)(m * 8) + a_off, a_hi, b_lo0 + b_off, b_hi, idesc,
                              (tap | kk) != 0 ? 1u : 0u);
            }
          }
#pragma unroll
          for (int m = 0; m < MT; ++m) umma_commit(&tfull[acc0 + m]);  // accumulators ready for the epilogue
          it += MT;
          umma_commit(&empty[stage]);  // window may be overwritten once these MMAs retire
        }
      }
    }
  } else if (warp < 8) {
    // ===================== epilogue =====================
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int h = warp >> 2;    // column half
    constexpr int HC = COUT / 2;      // columns per thread
    const int r = 32 * q + lane;      // accumulator row = TMEM lane
    const int g = r >> 3;             // feature column within the tile
    const int i = r & 7;              // time step within the tile
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int gc = 1 + kColTile * u + g;
      const int n = gc / p.cols;
      const int fp = gc - n * p.cols;
      const bool colvalid = (n < p.n_utts) && (fp >= 1) && (fp <= p.feats);

      if constexpr (Cfg::EPI == EPI_POOL_T) {
        static_assert(Cfg::EPI != EPI_POOL_T || HC == 32, "pool epilogue handles 32 columns per thread");
        const int odd = lane & 1;
        const int chbase = h * HC + odd * 16;  // 16 pooled channels owned by this lane
        const int RSo = Cfg::T / 2 + 2;
        for (int tt = 0; tt < Cfg::TILES; ++tt, ++it) {
          const int acc = it % NACC;
          mbar_wait(&tfull[acc], (it / NACC) & 1, 5);
          tc_fence_after();
          float v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + acc * COUT + h * HC, v);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
          // bias + ReLU (BN folded), then average the two time steps held by lanes (2k, 2k+1)
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c] + p.bias[h * HC + c], 0.0f);
          float o[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float send = odd ? v[c] : v[c + 16];
            const float mine = odd ? v[c + 16] : v[c];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
            o[c] = 0.5f * (mine + recv);
          }
          if (colvalid) {
            const int trow = 4 * tt + (i >> 1) + 1;  // pooled padded row: t' = 1+8tt+i (even lane) -> (t'+1)/2
            const long long rowoff = ((long long)gc * RSo + trow) * 8;
            const long long plane_elems = p.out_ncols * RSo * 8;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              uint16_t* dst = p.out + (long long)(chbase / 8 + k) * plane_elems + rowoff;
              st_global_v4(dst, pack_act2(o[8 * k + 0], o[8 * k + 1]), pack_act2(o[8 * k + 2], o[8 * k + 3]),
                           pack_act2(o[8 * k + 4], o[8 * k + 5]), pack_act2(o[8 * k + 6], o[8 * k + 7]));
            }
          }
        }
      } else {
        // time-sum of ReLU outputs kept in registers across the unit's tiles; no atomics, fixed order
        float sum[HC];
#pragma unroll
        for (int c = 0; c < HC; ++c) sum[c] = 0.0f;
        for (int tt = 0; tt < Cfg::TILES; ++tt, ++it) {
          const int acc = it % NACC;
          mbar_wait(&tfull[acc], (it / NACC) & 1, 5);
          tc_fence_after();
#pragma unroll
          for (int blk = 0; blk < HC / 32; ++blk) {
            float v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + acc * COUT + h * HC + blk * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) sum[blk * 32 + c] += fmaxf(v[c] + p.bias[h * HC + blk * 32 + c], 0.0f);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        // transpose-reduce over the 8 time lanes of a feature column: after the three steps lane
        // i holds the complete sums of HC/8 consecutive channels starting at (HC/8)*bitrev-free index.
        constexpr int W1 = HC / 2, W2 = HC / 4, W3 = HC / 8;
        {
          const bool up = (lane & 4) != 0;
#pragma unroll
          for (int c = 0; c < W1; ++c) {
            const float send = up ? sum[c] : sum[c + W1];
            const float keep = up ? sum[c + W1] : sum[c];
            sum[c] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
        }
        {
          const bool up = (lane & 2) != 0;
#pragma unroll
          for (int c = 0; c < W2; ++c) {
            const float send = up ? sum[c] : sum[c + W2];
            const float keep = up ? sum[c + W2] : sum[c];
            sum[c] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
        }
        {
          const bool up = (lane & 1) != 0;
#pragma unroll
          for (int c = 0; c < W3; ++c) {
            const float send = up ? sum[c] : sum[c + W3];
            const float keep = up ? sum[c + W3] : sum[c];
            sum[c] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
          }
        }
        if (colvalid) {
          const int cstart = h * HC + ((lane & 4) ? W1 : 0) + ((lane & 2) ? W2 : 0) + ((lane & 1) ? W3 : 0);
          float* dst = p.emb + ((long long)n * p.feats + (fp - 1)) * COUT + cstart;
#pragma unroll
          for (int c = 0; c < W3; c += 4)
            *reinterpret_cast<float4*>(dst + c) = make_float4(sum[c], sum[c + 1], sum[c + 2], sum[c + 3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
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

// Tensor map over one FT8 activation buffer: dim0 = (row, 8 channels) flattened and contiguous,
// dim1 = column, dim2 = plane; box = (wrows*8, 18, planes).
int make_act_tensor_map(CUtensorMap* out, const ActBuf& a, int wrows) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  DFS_REQUIRE(enc != nullptr, DFS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[3] = {(cuuint64_t)a.RS * 8, (cuuint64_t)a.ncols, (cuuint64_t)a.planes};
  cuuint64_t gstr[2] = {(cuuint64_t)a.RS * 16, (cuuint64_t)a.ncols * a.RS * 16};
  cuuint32_t box[3] = {(cuuint32_t)wrows * 8, (cuuint32_t)(kColTile + 2), (cuuint32_t)a.planes};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a.ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DFS_REQUIRE(r == CUDA_SUCCESS, DFS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return DFS_OK;
}

template <class Cfg>
static int launch_conv(const CUtensorMap& tmap, const ConvParams& p, int num_sms, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    DFS_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_tc_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_B));
    configured = true;
  }
  if (p.n_units <= 0) return DFS_OK;
  const int grid = p.n_units < num_sms * Cfg::OCC ? p.n_units : num_sms * Cfg::OCC;
  conv3x3_tc_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM_B, stream>>>(tmap, p);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

// CNN2D conv2: 32 -> 64 channels on 160 x 180, pooled to 80 rows.
using Conv2Cfg = ConvCfg<32, 64, 160, 2, 3, 4, EPI_POOL_T>;
// CNN2D conv3: 64 -> 128 channels on 80 x 180, summed over time.
using Conv3Cfg = ConvCfg<64, 128, 80, 2, 2, 4, EPI_MEAN_T>;

int conv2_tc_window_rows() { return Conv2Cfg::WROWS; }
int conv3_tc_window_rows() { return Conv3Cfg::WROWS; }

int launch_cnn2d_conv2_tc(const CUtensorMap& tmap_act1, const uint16_t* wpack, const float* bias, int n_utts, ActBuf act2,
                          int num_sms, cudaStream_t stream) {
  ConvParams p{};
  p.wpack = wpack;
  for (int i = 0; i < 64; ++i) p.bias[i] = bias[i];
  p.n_units = num_col_tiles(n_utts, kCols);
  p.n_utts = n_utts;
  p.cols = kCols;
  p.feats = kF;
  p.out = act2.ptr;
  p.out_ncols = act2.ncols;
  return launch_conv<Conv2Cfg>(tmap_act1, p, num_sms, stream);
}

int launch_cnn2d_conv3_tc(const CUtensorMap& tmap_act2, const uint16_t* wpack, const float* bias, int n_utts, float* emb,
                          int num_sms, cudaStream_t stream) {
  ConvParams p{};
  p.wpack = wpack;
  for (int i = 0; i < 128; ++i) p.bias[i] = bias[i];
  p.n_units = num_col_tiles(n_utts, kCols);
  p.n_utts = n_utts;
  p.cols = kCols;
  p.feats = kF;
  p.emb = emb;
  return launch_conv<Conv3Cfg>(tmap_act2, p, num_sms, stream);
}

}  // namespace dfs
