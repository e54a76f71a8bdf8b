// conv_tc.cu -- 3x3 convolution (+folded BN +ReLU +pool / +time-mean) as an implicit GEMM on the
// 5th-gen tensor cores: tcgen05.mma with fp32 accumulators in TMEM, operands staged by TMA.
//
// Replaces, on the scoring path, the reference's library calls
//   nn.Conv2d(32,64,3,p=1)+BatchNorm2d+ReLU+AvgPool2d((2,1))   /root/reference/src/model.py:21-24   (EPI_PAIR_POOL)
//   nn.Conv2d(64,128,3,p=1)+BatchNorm2d+ReLU, x.mean(dim=2)    /root/reference/src/model.py:27-29,37 (EPI_MEAN_T)
//
// GEMM view (see layout.cuh for the activation layouts):
//   one MMA tile  = 128 rows = 16 feature columns x 8 consecutive row indices of the input layout  (M = 128)
//   A (activations): SWIZZLE_NONE K-major smem descriptor straight into the TMA-loaded window; a tap is
//                    a compile-time constant added to the descriptor's start-address field
//   B (weights)    : BN-folded fp16, resident in shared memory for the whole kernel
//   D              : TMEM, NACC accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Two formulations, chosen by the measured single-CTA tcgen05 cost model (one MMA costs
// max(A-fetch 4 KB / ~48 B/clk, B-fetch / ~50 B/clk), i.e. ~88 cycles for any N <= 128):
//   MEAN (conv3): row = (feature column, time step); N = COUT = 128; K = 9 taps x 64 channels = 36 MMAs per tile.
//   PAIR (conv2): the input is stored with even/odd time steps in separate planes (FT8P), a row is a PAIR of
//                 output time steps (2j, 2j+1) and N = 2 x COUT = 128: columns [0,64) are the conv output at 2j,
//                 [64,128) at 2j+1.  K = 4 input time steps x 3 feature taps x 32 channels (one of the four time
//                 steps has zero weights for each half) = 24 MMAs per 256 conv outputs instead of 36 at N = 64,
//                 and the (2,1) average pool becomes an in-thread add of column c and column 64+c (no shuffles).
// Warp roles (352 threads): warps 0..7 = epilogue (TMEM lane quarter = warp%4, column half = warp/4),
// warp 8 = TMA producer, warp 9 = MMA issuer (one lane), warp 10 = TMEM allocator.
// Work unit = one column tile (16 feature columns, all rows); units are dealt round-robin to a persistent grid.
#include "common.cuh"
#include "kernels.h"
#include "layout.cuh"

namespace dfs {

enum { EPI_PAIR_POOL = 0, EPI_MEAN_T = 1 };

template <int CIN_, int COUT_, int ROWS_, int MT_, int NSTAGE_, int NACC_, int EPI_>
struct ConvCfg {
  static constexpr int CIN = CIN_, COUT = COUT_, ROWS = ROWS_, MT = MT_, NSTAGE = NSTAGE_, NACC = NACC_, EPI = EPI_;
  static constexpr bool PAIR = (EPI == EPI_PAIR_POOL);
  static constexpr int NG = PAIR ? 2 * COUT : COUT;    // GEMM N
  static constexpr int NTAP = PAIR ? 12 : 9;           // (input time step, feature tap) combinations
  static constexpr int CCH = CIN / 8;                  // 16-byte channel chunks
  static constexpr int KCH = PAIR ? 2 * CCH : CCH;     // planes of the input layout (PAIR: x2 time parities)
  static constexpr int WROWS = 8 * MT + 2;             // window rows incl. halo
  static constexpr int WCOLS = kColTile + 2;           // window columns (feature) incl. halo
  static constexpr int PLANE_B = WCOLS * WROWS * 16;   // bytes of one plane of the window
  static constexpr int WIN_B = KCH * PLANE_B;          // TMA transaction bytes per window
  static constexpr int WIN_B_AL = (WIN_B + 1023) & ~1023;
  static constexpr int WGT_B = NTAP * CIN * NG * 2;
  static constexpr int WGT_B_AL = (WGT_B + 1023) & ~1023;
  static constexpr int ST = ROWS / (8 * MT);           // windows (super-tiles) per unit
  static constexpr int TILES = ROWS / 8;               // MMA tiles per unit
  static constexpr int TMEM_COLS = NACC * NG;
  static constexpr int BAR_B = 256;
  static constexpr int SMEM_B = WGT_B_AL + NSTAGE * WIN_B_AL + BAR_B;
  static constexpr int THREADS = 352;
  // two CTAs per SM when shared memory and TMEM allow
  static constexpr int OCC = (SMEM_B <= 113 * 1024 && TMEM_COLS <= 256) ? 2 : 1;
  static_assert(ROWS % (8 * MT) == 0, "rows must be a multiple of the super-tile height");
  static_assert(NACC % MT == 0, "the accumulators of one window must be consecutive");
  static_assert(WROWS * 8 <= 256, "TMA box inner dimension limit");
  static_assert(TMEM_COLS == 32 || TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns");
  static_assert(NG % 64 == 0 && NG <= 256 && CIN % 16 == 0, "shape");
  static_assert(SMEM_B <= 227 * 1024, "shared memory budget");

  // byte offset of the A start address for (tap, 16-channel K step kk) relative to the tile's first row
  __host__ __device__ static constexpr int a_off(int tap, int kk) {
    if (PAIR) {
      // tap = r*3 + kw; input time step r in 0..3 relative to 2j-1: r=0 -> odd plane, row-1; r=1 -> even, row;
      // r=2 -> odd, row; r=3 -> even, row+1  (rows are pair indices; the window starts one row early)
      const int r = tap / 3, kw = tap % 3;
      const int par = (r == 0 || r == 2) ? 1 : 0;
      const int rowoff = (r == 0) ? 0 : (r == 3) ? 2 : 1;
      return (par * CCH + 2 * kk) * PLANE_B + (kw * WROWS + rowoff) * 16;
    }
    const int kh = tap / 3, kw = tap % 3;
    return (2 * kk) * PLANE_B + (kw * WROWS + kh) * 16;
  }
  __host__ __device__ static constexpr int b_off(int tap, int kk) { return ((tap * CCH + 2 * kk) * NG) * 16; }
};

struct ConvParams {
  const uint16_t* wpack;  // [NTAP][CIN/8][NG][8] fp16, BN folded
  float bias[128];        // folded bias per output channel (PAIR: already x0.5)
  int n_units;            // column tiles
  int n_utts;
  int cols;               // padded feature columns per utterance (F + 2)
  int feats;              // F
  // EPI_PAIR_POOL: pooled fp16 activations, FT8 layout with RS = ROWS + 2
  uint16_t* out;
  long long out_ncols;
  // EPI_MEAN_T: per-utterance time SUMS, [n][F][COUT] fp32 (the head applies 1/T)
  float* emb;
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::OCC)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ConvParams p) {
  constexpr int CIN = Cfg::CIN, COUT = Cfg::COUT, MT = Cfg::MT, NSTAGE = Cfg::NSTAGE, NACC = Cfg::NACC, NG = Cfg::NG;
  constexpr int WROWS = Cfg::WROWS, PLANE_B = Cfg::PLANE_B;
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
      constexpr uint32_t idesc = umma_idesc_f16(128, NG);
      // descriptor = (low word: start address >> 4 | LBO >> 4 << 16, high word: SBO >> 4 | version); taps and
      // K steps only move the start address, i.e. add a compile-time constant to the low word
      const uint64_t b_desc0 = umma_smem_desc(smem_u32(wsm), NG * 16, 128);
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
          // the MT tiles of a window are issued interleaved (tile index innermost)
          const int acc0 = it % NACC;  // NACC % MT == 0: the MT accumulators of a window are consecutive
#pragma unroll
          for (int m = 0; m < MT; ++m) mbar_wait(&tempty[acc0 + m], (((it + m) / NACC) & 1) ^ 1, 4);
          tc_fence_after();
#pragma unroll
          for (int tap = 0; tap < Cfg::NTAP; ++tap) {
#pragma unroll
            for (int kk = 0; kk < CIN / 16; ++kk) {
              const uint32_t a_off = (uint32_t)(Cfg::a_off(tap, kk) >> 4);
              const uint32_t b_off = (uint32_t)(Cfg::b_off(tap, kk) >> 4);
#pragma unroll
              for (int m = 0; m < MT; ++m)  // tile m = rows 8m.. of the window: +8 rows of 16 B
                umma_f16_lohi(tmem_base + (acc0 + m) * NG, a_lo_stage + (uint32_t)(m * 8) + a_off, a_hi, b_lo0 + b_off, b_hi, idesc,
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
    const int h = warp >> 2;          // output-channel half
    constexpr int HC = COUT / 2;      // output channels per thread
    const int r = 32 * q + lane;      // accumulator row = TMEM lane
    const int g = r >> 3;             // feature column within the tile
    const int i = r & 7;              // row within the tile
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int gc = 1 + kColTile * u + g;
      const int n = gc / p.cols;
      const int fp = gc - n * p.cols;
      const bool colvalid = (n < p.n_utts) && (fp >= 1) && (fp <= p.feats);

      if constexpr (Cfg::EPI == EPI_PAIR_POOL) {
        static_assert(Cfg::EPI != EPI_PAIR_POOL || HC == 32, "pair-pool epilogue handles 32 output channels per thread");
        constexpr int RSo = Cfg::ROWS + 2;
        const long long plane_elems = p.out_ncols * RSo * 8;
        for (int tt = 0; tt < Cfg::TILES; ++tt, ++it) {
          const int acc = it % NACC;
          mbar_wait(&tfull[acc], (it / NACC) & 1, 5);
          tc_fence_after();
          float a[32], b[32];  // conv outputs at time 2j (columns [0,COUT)) and 2j+1 (columns [COUT, 2 COUT))
          const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + acc * NG + h * HC;
          tmem_ld_32x32(taddr, a);
          tmem_ld_32x32(taddr + COUT, b);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
          // bias + ReLU on both time steps, sum = (2,1) average (0.5 folded into weights and bias)
          uint32_t pk[16];
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const float o0 = fmaxf(a[c] + p.bias[h * HC + c], 0.0f) + fmaxf(b[c] + p.bias[h * HC + c], 0.0f);
            const float o1 = fmaxf(a[c + 1] + p.bias[h * HC + c + 1], 0.0f) + fmaxf(b[c + 1] + p.bias[h * HC + c + 1], 0.0f);
            pk[c >> 1] = pack_act2(o0, o1);
          }
          if (colvalid) {
            uint16_t* dst = p.out + ((long long)gc * RSo + (8 * tt + i + 1)) * 8 + (long long)(4 * h) * plane_elems;
#pragma unroll
            for (int k = 0; k < 4; ++k) st_global_v4(dst + k * plane_elems, pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
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
            tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + acc * NG + h * HC + blk * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) sum[blk * 32 + c] += fmaxf(v[c] + p.bias[h * HC + blk * 32 + c], 0.0f);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        // transpose-reduce over the 8 time lanes of a feature column: after the three steps each lane
        // holds the complete sums of HC/8 consecutive channels.
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

// Tensor map over one FT8 / FT8P activation buffer: dim0 = (row, 8 channels) flattened and contiguous,
// dim1 = column, dim2 = plane; box = (wrows*8, 18, planes).
int make_act_tensor_map(CUtensorMap* out, const ActBuf& a, int wrows) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  DFS_REQUIRE(enc != nullptr, DFS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[3] = {(cuuint64_t)a.RS * 8, (cuuint64_t)a.ncols, (cuuint64_t)a.planes};
  cuuint64_t gstr[2] = {(cuuint64_t)a.RS * 16, (cuuint64_t)a.ncols * a.RS * 16};
  cuuint32_t box[3] = {(cuuint32_t)wrows * 8, (cuuint32_t)(kColTile + 2), (cuuint32_t)a.planes};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, a.ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
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

// CNN2D conv2: 32 -> 64 channels on 160 x 180 as 80 time PAIRS per column, pooled to 80 rows.
using Conv2Cfg = ConvCfg<32, 64, 80, 2, 3, 4, EPI_PAIR_POOL>;
// CNN2D conv3: 64 -> 128 channels on 80 x 180, summed over time.
using Conv3Cfg = ConvCfg<64, 128, 80, 2, 2, 4, EPI_MEAN_T>;

int conv2_tc_window_rows() { return Conv2Cfg::WROWS; }
int conv3_tc_window_rows() { return Conv3Cfg::WROWS; }

int launch_cnn2d_conv2_tc(const CUtensorMap& tmap_act1, const uint16_t* wpack, const float* bias_half, int n_utts, ActBuf act2,
                          int num_sms, cudaStream_t stream) {
  ConvParams p{};
  p.wpack = wpack;
  for (int i = 0; i < 64; ++i) p.bias[i] = bias_half[i];
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
