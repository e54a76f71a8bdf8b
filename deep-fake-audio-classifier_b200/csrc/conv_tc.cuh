// conv_tc.cuh -- the tcgen05 implicit-GEMM convolution template shared by the 2D-CNN (conv_tc.cu) and
// the convolutional autoencoder (cae_tc.cu).
//
// GEMM view (layouts: layout.cuh):
//   one MMA tile  = 128 rows = 16 feature columns x 8 consecutive row indices of the input layout  (M = 128)
//   A (activations): SWIZZLE_NONE K-major smem descriptor straight into the TMA-loaded window; a tap is a
//                    compile-time constant added to the descriptor's start-address field
//   B (weights)    : BN-folded fp16, resident in shared memory for the whole kernel
//   D              : TMEM, NACC accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// MODE_3X3  : row = (feature column, time step); N = COUT; K = 9 taps x CIN.
// MODE_PAIR : the input is stored with even/odd time steps in separate planes (FT8P), a row is a PAIR of output
//             time steps (2j, 2j+1), N = 2 x COUT (columns [0,COUT) = output at 2j, [COUT,2 COUT) = at 2j+1),
//             K = 4 input time steps x 3 feature taps x CIN (zero weights where a time step does not reach an
//             output).  Chosen when COUT = 64: the measured single-CTA MMA cost is ~88 cycles for any N <= 128
//             (DESIGN.md §4), so N = 128 with 4/3 of the MACs beats N = 64; the time pool becomes in-thread.
// MODE_1X1  : one tap, no halo: the k=2,s=2 transposed convolutions of the CAE decoder are GEMMs over positions
//             with N = (output quadrant, COUT) and a pixel-shuffle epilogue.
// MODE_3X3S : 3x3 with SWAPPED operand roles: A = weights (M = COUT = 128 rows), B = activations with N = 256 positions
//             (32 feature columns x 8 time steps).  Per the cost model N = 256 reaches 83 % of the tensor pipe instead
//             of 73 % at N = 128, the accumulator row is an output channel, so the time sum is an in-thread add over
//             8 consecutive columns and the [n][f][c] store is 128 contiguous bytes per warp.
// MODE_3X1  : three taps along the row (time) axis only, no column halo: the Conv1d(k=3) layers of the 1D-CNN with one
//             "feature column" per utterance (cols = 1: a tile is 16 utterances x 8 time steps).
// MODE_5X1  : the same with five taps (Conv1d(k=5, p=2) of the StatsPool detector); the two halo rows per side that the
//             padded layout does not store come from the TMA's out-of-bounds zero fill.
// KSPLIT    : the CIN/8 channel planes of a window are loaded as KSPLIT separate pipeline stages ("pieces"),
//             which bounds shared memory for CIN >= 128.
// blockIdx.y: output-channel / quadrant group (weights, bias and output placement are offset per group).
//
// Warp roles (352 threads): warps 0..7 = epilogue (TMEM lane quarter = warp%4, column half = warp/4),
// warp 8 = TMA producer, warp 9 = MMA issuer (one lane), warp 10 = TMEM allocator.
// Work unit = one column tile (16 feature columns, all rows); units are dealt round-robin to a persistent grid.
#pragma once
#include "common.cuh"
#include "kernels.h"
#include "layout.cuh"

namespace dfs {

enum { MODE_3X3 = 0, MODE_PAIR = 1, MODE_1X1 = 2, MODE_3X1 = 3, MODE_3X3S = 4, MODE_5X1 = 5 };
enum {
  EPI_PAIR_POOL = 0,    // PAIR: relu both time steps, add (time pool), store FT8                    (CNN2D conv2)
  EPI_MEAN_T = 1,       // 3x3 : relu, sum over all rows of the unit, store [n][F][COUT] fp32         (CNN2D conv3)
  EPI_PAIR_POOL_F = 2,  // PAIR: time pool in-thread + feature pool with lane^8, store FT8            (CAE enc2)
  EPI_POOL_TF = 3,      // 3x3 : relu, 2x2 pool with lane^1 (time) and lane^8 (feature), store FT8    (CAE enc4; enc3 with enc3_swap = 0)
  EPI_SHUFFLE = 4,      // 1x1 : relu, pixel-shuffle store of the quadrant(s) held in the columns      (CAE dec1-3)
  EPI_RELU = 5,         // any : relu, store FT8 at the same position                                   (CNN1D layers 1, 2)
  EPI_MEAN_T_SWAP = 6,  // 3x3S: lanes = output channels, columns = positions; relu, time sum in-thread   (CNN2D conv3)
  EPI_SHUFFLE_MSE = 7,  // 1x1 : CAE dec3 with the final ConvTranspose2d(32,1) and the squared error against the input fused in:
                        //        neither d3 nor the reconstruction is written; one partial sum per 16-column unit   (CAE dec3+final)
  EPI_GELU = 8,         // any : exact (erf) GELU, store FT8 at the same position, output-channel groups          (StatsPool detector)
  EPI_SHUFFLE_ROWS = 9, // 1x1 : pixel shuffle with both row offsets (a = 0, 1) of a 32-channel block in one thread; the a = 1 row moves
                        //        one lane up so that every thread writes whole 32-byte sectors (256-bit stores)        (CAE dec1, dec2)
  EPI_POOL_TF_SWAP = 10 // 3x3S: lanes = output channels, columns = positions: relu, 2x2 pool entirely in-thread, 2-byte stores    (CAE enc3)
};

// CTA2 = 1: the kernel runs as CTA pairs (cluster of 2, tcgen05 cta_group::2): one MMA covers the two units of a pair (M = 256) and
//           each CTA keeps only NG/2 of the NG weight rows in ITS shared memory.  Measured (profiles/r01f_umma_bench_pairs.txt): a
//           pair MMA costs the cycles of a single-CTA MMA of the same N, i.e. per SM nothing is gained unless shared memory had
//           forced N < 128 on one CTA (CAE enc4, StatsPool layer 1: 147 / 123 KB of weights per 64 output channels).
// SPLIT = 1: the "split" precision (option precision = 2): every activation and every weight is carried as fp16 value + fp16 rounding
//           residual (x = hi + lo up to 2^-22 |x|) and a product is three MMAs into the same fp32 accumulator, A_hi W_hi + A_hi W_lo +
//           A_lo W_hi (the dropped A_lo W_lo term is 2^-22 of the product).  The input buffer holds the residual planes after the value
//           planes (2 x KCH planes); a window is loaded as two pieces (value planes, residual planes) and the weight image of a group is
//           [value image | residual image].  Epilogues that store activations write value and residual planes.
template <int MODE_, int CIN_, int COUT_, int NG_, int ROWS_, int MT_, int NSTAGE_, int NACC_, int KSPLIT_, int EPI_, int CTA2_ = 0, int SPLIT_ = 0>
struct ConvCfg {
  static constexpr int MODE = MODE_, CIN = CIN_, COUT = COUT_, NG = NG_, ROWS = ROWS_, MT = MT_, NSTAGE = NSTAGE_, NACC = NACC_,
                       KSPLIT = KSPLIT_, EPI = EPI_, CTA2 = CTA2_, SPLIT = SPLIT_;
  static constexpr bool PAIR = (MODE == MODE_PAIR);
  static constexpr bool SWAP = (MODE == MODE_3X3S);
  static constexpr int CT = SWAP ? 32 : kColTile;      // feature columns per tile
  static constexpr int HALO = (MODE == MODE_1X1) ? 0 : (MODE == MODE_5X1 ? 2 : 1);   // row halo
  static constexpr int HALO_C = (MODE == MODE_1X1 || MODE == MODE_3X1 || MODE == MODE_5X1) ? 0 : 1;  // column halo
  static constexpr int NTAP = PAIR ? 12 : (MODE == MODE_1X1 ? 1 : (MODE == MODE_3X1 ? 3 : (MODE == MODE_5X1 ? 5 : 9)));
  static constexpr int CCH = CIN / 8;                  // 16-byte channel chunks
  static constexpr int KCH = PAIR ? 2 * CCH : CCH;     // planes of the input layout (PAIR: x2 time parities)
  static constexpr int NPIECE = SPLIT ? 2 : KSPLIT;    // pipeline pieces per window (SPLIT: value planes, residual planes)
  static constexpr int PPL = SPLIT ? KCH : KCH / KSPLIT;   // planes per piece
  static constexpr int CPP = SPLIT ? CCH : CCH / KSPLIT;   // channel chunks per piece
  static constexpr int WROWS = 8 * MT + 2 * HALO;      // window rows incl. halo
  static constexpr int WCOLS = CT + 2 * HALO_C;        // window columns (feature) incl. halo
  static constexpr int PLANE_B = WCOLS * WROWS * 16;   // bytes of one plane of the window
  static constexpr int WIN_B = PPL * PLANE_B;          // TMA transaction bytes per piece
  static constexpr int WIN_B_AL = (WIN_B + 1023) & ~1023;
  static constexpr int WROWS_OUT = SWAP ? COUT : (CTA2 ? NG / 2 : NG);   // rows of the weight operand image (per CTA)
  static constexpr int WGT_TERM_B = NTAP * CIN * WROWS_OUT * 2;   // one weight image
  static constexpr int WGT_B = (SPLIT ? 2 : 1) * WGT_TERM_B;      // per output group (SPLIT: value image | residual image)
  static constexpr int WGT_B_AL = (WGT_B + 1023) & ~1023;
  static constexpr int ST = ROWS / (8 * MT);           // windows (super-tiles) per unit
  static constexpr int TILES = ROWS / 8;               // MMA tiles per unit
  static constexpr int TMEM_COLS = NACC * NG;
  static constexpr int BAR_B = 256;
  static constexpr int SMEM_B = WGT_B_AL + NSTAGE * WIN_B_AL + BAR_B;
  static constexpr int THREADS = 352;
  // two CTAs per SM when shared memory and TMEM allow
  static constexpr int OCC = (!CTA2 && SMEM_B <= 113 * 1024 && TMEM_COLS <= 256) ? 2 : 1;
  static_assert(!CTA2 || (!SWAP && NG % 32 == 0), "CTA pairs: the activation window is the A operand");
  static_assert(ROWS % (8 * MT) == 0, "rows must be a multiple of the super-tile height");
  static_assert(NACC % MT == 0, "the accumulators of one window must be consecutive");
  static_assert(WROWS * 8 <= 256, "TMA box inner dimension limit");
  static_assert(TMEM_COLS == 32 || TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns");
  static_assert(NG % 64 == 0 && NG <= 256 && CIN % 16 == 0, "shape");
  static_assert(!PAIR || KSPLIT == 1, "PAIR mode loads both parities in one piece");
  static_assert(!SPLIT || (KSPLIT == 1 && !SWAP), "split precision: the two pieces are the value and the residual planes");
  static_assert(!SWAP || (MT == 1 && NG == 256 && COUT == 128), "swapped mode: one 128 x 256 tile per window");
  static_assert(CCH % KSPLIT == 0 && CPP % 2 == 0, "a piece must hold whole K=16 steps");
  static_assert(WGT_TERM_B % 16 == 0, "descriptor start addresses are in 16-byte units");
  static_assert(SMEM_B <= 227 * 1024, "shared memory budget");

  // byte offset of the A start address for (tap, K step kk of the piece) relative to the tile's first row in the piece window
  __host__ __device__ static constexpr int a_off(int tap, int kk) {
    if (MODE == MODE_PAIR) {
      // tap = r*3 + kw; input time step r in 0..3 relative to 2j-1: r=0 -> odd plane, row-1; r=1 -> even, row;
      // r=2 -> odd, row; r=3 -> even, row+1  (rows are pair indices; the window starts one row early)
      const int r = tap / 3, kw = tap % 3;
      const int par = (r == 0 || r == 2) ? 1 : 0;
      const int rowoff = (r == 0) ? 0 : (r == 3) ? 2 : 1;
      return (par * CCH + 2 * kk) * PLANE_B + (kw * WROWS + rowoff) * 16;
    }
    if (MODE == MODE_1X1) return (2 * kk) * PLANE_B;
    if (MODE == MODE_3X1 || MODE == MODE_5X1) return (2 * kk) * PLANE_B + tap * 16;
    const int kh = tap / 3, kw = tap % 3;
    return (2 * kk) * PLANE_B + (kw * WROWS + kh) * 16;
  }
  // byte offset of the B start address for (tap, piece, K step kk)
  __host__ __device__ static constexpr int b_off(int tap, int piece, int kk) { return ((tap * CCH + piece * CPP + 2 * kk) * WROWS_OUT) * 16; }
};

struct ConvParams {
  const uint16_t* wpack;  // [group][NTAP][CIN/8][NG][8] fp16, BN folded
  float bias[256];        // folded bias per OUTPUT CHANNEL (all groups), pre-scaled like the weights
  int n_units;            // column tiles
  int n_utts;
  int cols;               // padded feature columns per utterance of the INPUT layout
  int feats;              // valid feature columns of the input
  int rows_valid;         // valid rows of the input (<= ROWS; rows beyond are padding whose outputs are dropped)
  // fp16 output (FT8): geometry of the OUTPUT layout
  uint16_t* out;
  long long out_ncols;    // allocated columns per plane
  int out_rs;             // rows per column
  int out_cols;           // padded feature columns per utterance
  int out_feats;          // valid output feature columns (pooled outputs beyond are dropped)
  // EPI_MEAN_T: per-utterance time SUMS, [n][F][COUT] fp32 (the head applies 1/T)
  float* emb;
  // EPI_SHUFFLE_MSE: the scorer's input (element (n,t,f) at x[n*xsn + t*xst + f*xsf]), the optional normaliser and the
  // per-unit partial sums; the final layer's weights ride in bias[32..159] ([ci*4 + a*2+b]) and its bias in bias[160]
  const float* x;
  long long xsn, xst, xsf;
  const float* norm_mean;
  const float* norm_sd;
  float* partial;
  int x_vec4;             // 1 = x rows are 16-byte aligned with the feature axis contiguous: float4 loads
  float inv_scale;        // SPLIT: 1 / (power-of-two scale of the weight images); the accumulator is multiplied by it
};

// ---- epilogue helpers -----------------------------------------------------------------------------
// accumulator drained: tell the MMA issuer (CTA pairs: the leader's barrier counts the 8 epilogue warps of both CTAs)
template <int CTA2>
__device__ __forceinline__ void acc_release(uint64_t* tempty_bar) {
  if constexpr (CTA2) mbar_arrive_cluster(mapa_u32(smem_u32(tempty_bar), 0));
  else mbar_arrive(tempty_bar);
}
template <int NCH>
__device__ __forceinline__ void store_chunks(uint16_t* dst, long long plane_elems, const uint32_t* pk) {
#pragma unroll
  for (int k = 0; k < NCH; ++k) st_global_v4(dst + k * plane_elems, pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::OCC)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ConvParams p) {
  constexpr int COUT = Cfg::COUT, MT = Cfg::MT, NSTAGE = Cfg::NSTAGE, NACC = Cfg::NACC, NG = Cfg::NG;
  constexpr int WROWS = Cfg::WROWS, PLANE_B = Cfg::PLANE_B, HALO = Cfg::HALO, HALO_C = Cfg::HALO_C;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* wsm = smem;
  uint8_t* win0 = smem + Cfg::WGT_B_AL;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::WGT_B_AL + NSTAGE * Cfg::WIN_B_AL);
  uint64_t* full = bars;                    // [NSTAGE]  TMA -> MMA
  uint64_t* empty = bars + NSTAGE;          // [NSTAGE]  MMA -> TMA
  uint64_t* tfull = bars + 2 * NSTAGE;      // [NACC]    MMA -> epilogue
  uint64_t* tempty = tfull + NACC;          // [NACC]    epilogue -> MMA
  uint64_t* wbar = tempty + NACC;           // weights resident
  uint64_t* pwbar = wbar + 1;               // CTA pairs: the peer's weights are resident (leader's copy is the one waited on)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pwbar + 1);
  static_assert((2 * NSTAGE + 2 * NACC + 2) * 8 + 4 <= 160, "barrier page");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.y;               // output-channel / quadrant group
  // CTA pairs: rank 0 (leader) and rank 1 work on the units 2j and 2j+1 in lockstep; a unit index past n_units is all padding
  const int rank = Cfg::CTA2 ? (int)cluster_ctarank() : 0;
  const int u_first = Cfg::CTA2 ? (int)(blockIdx.x & ~1u) + rank : (int)blockIdx.x;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap);
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], Cfg::CTA2 ? 16 : 8); }
    mbar_init(wbar, 1);
    mbar_init(pwbar, 1);
    fence_mbar_init();
  }
  if (warp == 10) {
    if constexpr (Cfg::CTA2) {
      tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (Cfg::CTA2) cluster_sync_all();   // the peer's barriers must be initialised before anything is signalled across
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(wbar, Cfg::WGT_B);
      constexpr int PIECE = 16384;
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)(Cfg::CTA2 ? 2 * grp + rank : grp) * Cfg::WGT_B;
      for (int off = 0; off < Cfg::WGT_B; off += PIECE) {
        const int bytes = (Cfg::WGT_B - off) < PIECE ? (Cfg::WGT_B - off) : PIECE;
        bulk_g2s(wsm + off, wsrc + off, bytes, wbar);
      }
      uint32_t ws = 0;
      for (int u = u_first; u - rank < p.n_units; u += gridDim.x) {
        for (int st = 0; st < Cfg::ST; ++st) {
#pragma unroll 1
          for (int pc = 0; pc < Cfg::NPIECE; ++pc, ++ws) {
            const int stage = ws % NSTAGE;
            mbar_wait(&empty[stage], ((ws / NSTAGE) & 1) ^ 1, 1);
            if constexpr (Cfg::CTA2) {   // both windows of the pair are counted on the leader's barrier
              if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * Cfg::WIN_B);
              tma_load_3d_pair(win0 + stage * Cfg::WIN_B_AL, &tmap, (1 + 8 * MT * st - HALO) * 8, 1 + u * Cfg::CT - HALO_C, pc * Cfg::PPL,
                               mapa_u32(smem_u32(&full[stage]), 0));
            } else {
              mbar_arrive_expect_tx(&full[stage], Cfg::WIN_B);
              tma_load_3d(win0 + stage * Cfg::WIN_B_AL, &tmap, (1 + 8 * MT * st - HALO) * 8, 1 + u * Cfg::CT - HALO_C, pc * Cfg::PPL,
                          &full[stage]);
            }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (Cfg::CTA2 && rank == 1) {
      if (elect_one_sync()) {   // the peer only reports its weights
        mbar_wait(wbar, 0, 2);
        mbar_arrive_cluster(mapa_u32(smem_u32(pwbar), 0));
      }
    } else if (elect_one_sync()) {   // elect_one_sync(), not `lane == 0`: under a lane test the compiler cannot tell that a single thread is active and wraps EVERY tcgen05.mma / TMA issue in an ELECT + BRA.U.ANY serialisation loop (measured: ~100 instead of 64 cycles per N = 128 MMA)
      constexpr uint32_t idesc = umma_idesc_f16(Cfg::CTA2 ? 256 : 128, NG);
      [[maybe_unused]] constexpr uint32_t idesc_half = umma_idesc_f16(128, COUT < 16 ? 16 : COUT);   // PAIR: N = COUT for the half-width taps
      // descriptor = (low word: start address >> 4 | LBO >> 4 << 16, high word: SBO >> 4 | version); taps and
      // K steps only move the start address, i.e. add a compile-time constant to the low word
      const uint64_t b_desc0 = umma_smem_desc(smem_u32(wsm), Cfg::WROWS_OUT * 16, 128);
      const uint32_t b_lo0 = (uint32_t)b_desc0, b_hi = (uint32_t)(b_desc0 >> 32);
      const uint64_t a_desc0 = umma_smem_desc(smem_u32(win0), PLANE_B, WROWS * 16);
      const uint32_t a_lo0 = (uint32_t)a_desc0, a_hi = (uint32_t)(a_desc0 >> 32);
      mbar_wait(wbar, 0, 2);
      if constexpr (Cfg::CTA2) mbar_wait(pwbar, 0, 6);
      uint32_t ws = 0, it = 0;
      for (int u = u_first; u < p.n_units; u += gridDim.x) {
        for (int st = 0; st < Cfg::ST; ++st) {
          const int acc0 = it % NACC;  // NACC % MT == 0: the MT accumulators of a window are consecutive
#pragma unroll
          for (int pc = 0; pc < Cfg::NPIECE; ++pc, ++ws) {
            const int stage = ws % NSTAGE;
            mbar_wait(&full[stage], (ws / NSTAGE) & 1, 3);
            if (pc == 0) {
#pragma unroll
              for (int m = 0; m < MT; ++m) mbar_wait(&tempty[acc0 + m], (((it + m) / NACC) & 1) ^ 1, 4);
            }
            tc_fence_after();
            const uint32_t a_lo_stage = a_lo0 + (uint32_t)(stage * (Cfg::WIN_B_AL >> 4));
            // the MT tiles of a window are issued interleaved (tile index innermost)
            // SPLIT: the value planes (piece 0) meet the value and the residual image of the weights, the residual planes (piece 1) the value image
            const int nterm = (Cfg::SPLIT && pc == 0) ? 2 : 1;
#pragma unroll
            for (int term = 0; term < nterm; ++term) {
#pragma unroll
              for (int tq = 0; tq < Cfg::NTAP; ++tq) {
                // PAIR on a single CTA: the taps of input time step r = 0 reach only the first output of the pair (columns [0, COUT)) and
                // those of r = 3 only the second one (columns [COUT, 2 COUT)) -- the other half of their weight rows is zeros.  They are
                // issued as N = COUT MMAs into that half of the accumulator (48 instead of 64 cycles each, none of the zero MACs), after
                // the full-width taps r = 1, 2 so that the first MMA of a tile initialises all columns.
                constexpr bool HALF_TAPS = Cfg::PAIR && !Cfg::CTA2 && !Cfg::SPLIT;
                const int tap = HALF_TAPS ? (tq < 6 ? tq + 3 : (tq < 9 ? tq - 6 : tq)) : tq;
                const bool half = HALF_TAPS && (tap < 3 || tap >= 9);
                const uint32_t half_col = (HALF_TAPS && tap >= 9) ? (uint32_t)COUT : 0u;   // accumulator column = weight row offset of the r = 3 taps
#pragma unroll
                for (int kk = 0; kk < Cfg::CPP / 2; ++kk) {
                  const uint32_t a_off = (uint32_t)(Cfg::a_off(tap, kk) >> 4);
                  const uint32_t b_off = (uint32_t)((Cfg::SPLIT ? term * Cfg::WGT_TERM_B + Cfg::b_off(tap, 0, kk) : Cfg::b_off(tap, pc, kk)) >> 4) + half_col;
                  const uint32_t accum = (pc | term | tq | kk) != 0 ? 1u : 0u;
#pragma unroll
                  for (int m = 0; m < MT; ++m) {  // tile m = rows 8m.. of the window: +8 rows of 16 B
                    if constexpr (Cfg::SWAP)  // weights are the A (M) operand, the activation window is the B (N = 256) operand
                      umma_f16_lohi(tmem_base + (acc0 + m) * NG, b_lo0 + b_off, b_hi, a_lo_stage + a_off, a_hi, idesc, accum);
                    else if constexpr (Cfg::CTA2)
                      umma_f16_lohi_pair(tmem_base + (acc0 + m) * NG, a_lo_stage + (uint32_t)(m * 8) + a_off, a_hi, b_lo0 + b_off, b_hi, idesc, accum);
                    else
                      umma_f16_lohi(tmem_base + (acc0 + m) * NG + half_col, a_lo_stage + (uint32_t)(m * 8) + a_off, a_hi, b_lo0 + b_off, b_hi,
                                    half ? idesc_half : idesc, accum);
                  }
                }
              }
            }
            if (pc == Cfg::NPIECE - 1) {
#pragma unroll
              for (int m = 0; m < MT; ++m) {  // accumulators ready for the epilogue
                if constexpr (Cfg::CTA2) umma_commit_pair(&tfull[acc0 + m]);
                else umma_commit(&tfull[acc0 + m]);
              }
            }
            // window may be overwritten once these MMAs retire
            if constexpr (Cfg::CTA2) umma_commit_pair(&empty[stage]);
            else umma_commit(&empty[stage]);
          }
          it += MT;
        }
      }
    }
  } else if (warp < 8) {
    // ===================== epilogue =====================
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int h = warp >> 2;          // column half
    const int r = 32 * q + lane;      // accumulator row = TMEM lane
    const int g = r >> 3;             // feature column within the tile
    const int i = r & 7;              // row within the tile
    const float* bias = p.bias;  // param space: uniform constant-bank reads
    uint32_t it = 0;
    [[maybe_unused]] uint32_t unit_seq = 0;
    for (int u = u_first; u - rank < p.n_units; u += gridDim.x) {
      const int gc = 1 + Cfg::CT * u + g;
      // padded layouts: column gc = n*cols + f' (f' = 0 and cols-1 are zero pads); cols == 1: one column per utterance at gc = 1 + n
      const int n = (p.cols == 1) ? gc - 1 : gc / p.cols;
      const int fp = (p.cols == 1) ? 1 : gc - n * p.cols;
      const bool colvalid = (n < p.n_utts) && (fp >= 1) && (fp <= p.feats);

      if constexpr (Cfg::EPI == EPI_PAIR_POOL || Cfg::EPI == EPI_PAIR_POOL_F) {
        constexpr int HC = COUT / 2;  // output channels per thread
        const long long plane_elems = p.out_ncols * p.out_rs * 8;
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
          if (lane == 0) acc_release<Cfg::CTA2>(&tempty[acc]);
          // bias + ReLU on both time steps, sum = time pool (the pool's 1/2 or 1/4 is folded into weights and bias)
          float o[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            if constexpr (Cfg::SPLIT) o[c] = relu_nan(fmaf(a[c], p.inv_scale, bias[h * HC + c])) + relu_nan(fmaf(b[c], p.inv_scale, bias[h * HC + c]));
            else o[c] = relu_nan(a[c] + bias[h * HC + c]) + relu_nan(b[c] + bias[h * HC + c]);
          }
          const int row_out = 8 * tt + i + 1;  // pair index + 1 = padded output row
          if constexpr (Cfg::EPI == EPI_PAIR_POOL) {
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 32; c += 2) pk[c >> 1] = pack_act2(o[c], o[c + 1]);
            if (colvalid) {
              uint16_t* dst = p.out + ((long long)gc * p.out_rs + row_out) * 8 + (long long)(4 * h) * plane_elems;
              store_chunks<4>(dst, plane_elems, pk);
              if constexpr (Cfg::SPLIT) {   // residual planes follow the COUT / 8 value planes
                uint32_t pr[16];
#pragma unroll
                for (int c = 0; c < 32; c += 2) pr[c >> 1] = pack_act2_residual(o[c], o[c + 1], pk[c >> 1]);
                store_chunks<4>(dst + (long long)(COUT / 8) * plane_elems, plane_elems, pr);
              }
            }
          } else {
            // feature pool: columns g (odd f') and g+1 live in lanes l and l^8; each keeps 16 of the 32 channels
            const int oddg = g & 1;
            uint32_t pk[8];
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
              float v[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float send = oddg ? o[c + e] : o[c + e + 16];
                const float mine = oddg ? o[c + e + 16] : o[c + e];
                v[e] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
              }
              pk[c >> 1] = pack_act2(v[0], v[1]);
            }
            const int fo = (fp - 1) >> 1;  // pooled feature index of the pair (same for both lanes)
            const bool ok = (n < p.n_utts) && (fp >= 1) && (fo < p.out_feats);
            if (ok) {
              const long long gco = (long long)n * p.out_cols + fo + 1;
              uint16_t* dst = p.out + (gco * p.out_rs + row_out) * 8 + (long long)(4 * h + 2 * oddg) * plane_elems;
              store_chunks<2>(dst, plane_elems, pk);
            }
          }
        }
      } else if constexpr (Cfg::EPI == EPI_MEAN_T) {
        constexpr int HC = COUT / 2;
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
            if (8 * tt + i + 1 <= p.rows_valid) {  // rows beyond the valid range are padding (1D-CNN: 321 of 328)
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                if constexpr (Cfg::SPLIT) sum[blk * 32 + c] += relu_nan(fmaf(v[c], p.inv_scale, bias[h * HC + blk * 32 + c]));
                else sum[blk * 32 + c] += relu_nan(v[c] + bias[h * HC + blk * 32 + c]);
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) acc_release<Cfg::CTA2>(&tempty[acc]);
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
      } else if constexpr (Cfg::EPI == EPI_MEAN_T_SWAP) {
        // accumulator row (TMEM lane) = output channel, column j = 8*(feature column) + time step; this thread sums the
        // ReLU outputs over time for the 16 feature columns of its column half, across all tiles of the unit
        const int ch = 32 * q + lane;
        const float bc = bias[ch];
        float tsum[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) tsum[c] = 0.0f;
        for (int tt = 0; tt < Cfg::TILES; ++tt, ++it) {
          const int acc = it % NACC;
          mbar_wait(&tfull[acc], (it / NACC) & 1, 5);
          tc_fence_after();
#pragma unroll
          for (int blk = 0; blk < 4; ++blk) {
            float v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + acc * NG + h * 128 + blk * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) tsum[blk * 4 + (k >> 3)] += relu_nan(v[k] + bc);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) acc_release<Cfg::CTA2>(&tempty[acc]);
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int gcc = 1 + Cfg::CT * u + h * 16 + c;
          const int nn = gcc / p.cols;
          const int fpp = gcc - nn * p.cols;
          if (nn < p.n_utts && fpp >= 1 && fpp <= p.feats) p.emb[((long long)nn * p.feats + (fpp - 1)) * COUT + ch] = tsum[c];
        }
      } else if constexpr (Cfg::EPI == EPI_POOL_TF_SWAP) {
        // Swapped roles (weights = A, 256 positions = N): accumulator row (TMEM lane) = output channel, column j = 8 * (feature
        // column of the 32-column tile) + (row of the 8-row tile).  Both partners of the 2x2 average pool are columns of the same
        // thread: feature pairs (even tile column, +1) -- the tile starts at an odd padded column and utterances hold an even
        // number of columns, so a pair never straddles utterances -- and row pairs (even row, +1).  The 1/4 is folded into
        // weights and bias.  One pooled value = one fp16 of the FT8 chunk of its channel: 2-byte stores, 8 consecutive lanes
        // fill one 16-byte chunk.
        static_assert(Cfg::EPI != EPI_POOL_TF_SWAP || (Cfg::SWAP && COUT == 128 && NG == 256), "swapped pool epilogue shape");
        const int ch = 32 * q + lane;
        const float bc = bias[ch];
        const long long plane_elems = p.out_ncols * p.out_rs * 8;
        uint16_t* chbase = p.out + (long long)(ch >> 3) * plane_elems + (ch & 7);
        // the 8 feature pairs of this thread's column half: output column offset (elements) or -1 when invalid; fixed per unit
        long long coloff[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int gcc = 1 + Cfg::CT * u + h * 16 + 2 * k;      // padded input column of the pair's first member (odd f')
          const int nn = gcc / p.cols;
          const int fpp = gcc - nn * p.cols;
          const int fo = (fpp - 1) >> 1;
          const bool ok = (nn < p.n_utts) && (fpp >= 1) && (fpp + 1 <= p.feats) && (fo < p.out_feats);
          coloff[k] = ok ? ((long long)nn * p.out_cols + fo + 1) * p.out_rs * 8 : -1;
        }
        for (int tt = 0; tt < Cfg::TILES; ++tt, ++it) {
          const int acc = it % NACC;
          mbar_wait(&tfull[acc], (it / NACC) & 1, 5);
          tc_fence_after();
#pragma unroll
          for (int blk = 0; blk < 4; ++blk) {   // 32 columns = 4 tile columns x 8 rows = 2 feature pairs x 4 row pairs
            float v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + acc * NG + h * 128 + blk * 32, v);
            tmem_ld_wait();
            if (blk == 3) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) acc_release<Cfg::CTA2>(&tempty[acc]);
            }
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = relu_nan(v[k] + bc);
#pragma unroll
            for (int fpair = 0; fpair < 2; ++fpair) {
              const long long co = coloff[2 * blk + fpair];
#pragma unroll
              for (int rp = 0; rp < 4; ++rp) {
                const float s = (v[16 * fpair + 2 * rp] + v[16 * fpair + 2 * rp + 1]) + (v[16 * fpair + 8 + 2 * rp] + v[16 * fpair + 8 + 2 * rp + 1]);
                const int to = 4 * tt + rp;                          // pooled row of input rows 8 tt + 2 rp, +1
                if (co >= 0 && 8 * tt + 2 * rp + 2 <= p.rows_valid) {
                  const __half hv = __ushort_as_half((unsigned short)(pack_act2(s, 0.0f) & 0xffffu));
                  chbase[co + (long long)(to + 1) * 8] = __half_as_ushort(hv);
                }
              }
            }
          }
        }
      } else if constexpr (Cfg::EPI == EPI_POOL_TF) {
        // relu, then 2x2 average pool: time partner = lane^1, feature partner = lane^8 (1/4 folded into weights/bias).
        // Each exchange halves the channels a lane keeps: HC -> HC/2 -> HC/4.
        constexpr int HC = COUT / 2, Q1 = HC / 2, Q2 = HC / 4;
        static_assert(Cfg::EPI != EPI_POOL_TF || (HC % 32 == 0 && Q2 % 8 == 0), "pool epilogue channel split");
        const long long plane_elems = p.out_ncols * p.out_rs * 8;
        const int oddi = i & 1, oddg = g & 1;
        for (int tt = 0; tt < Cfg::TILES; ++tt, ++it) {
          const int acc = it % NACC;
          mbar_wait(&tfull[acc], (it / NACC) & 1, 5);
          tc_fence_after();
          float v[HC];
#pragma unroll
          for (int blk = 0; blk < HC / 32; ++blk) tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + acc * NG + h * HC + blk * 32, v + blk * 32);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) acc_release<Cfg::CTA2>(&tempty[acc]);
#pragma unroll
          for (int c = 0; c < HC; ++c) v[c] = relu_nan(v[c] + bias[grp * COUT + h * HC + c]);
#pragma unroll
          for (int c = 0; c < Q1; ++c) {
            const float send = oddi ? v[c] : v[c + Q1];
            const float mine = oddi ? v[c + Q1] : v[c];
            v[c] = mine + __shfl_xor_sync(0xffffffffu, send, 1);
          }
          uint32_t pk[Q2 / 2];
#pragma unroll
          for (int c = 0; c < Q2; c += 2) {
            float w2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float send = oddg ? v[c + e] : v[c + e + Q2];
              const float mine = oddg ? v[c + e + Q2] : v[c + e];
              w2[e] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            pk[c >> 1] = pack_act2(w2[0], w2[1]);
          }
          const int tp = 1 + 8 * tt + i;           // padded input row
          const int to = (tp - 1) >> 1;            // pooled row (same for both time lanes)
          const int fo = (fp - 1) >> 1;
          const bool ok = (n < p.n_utts) && (fp >= 1) && (fo < p.out_feats) && (tp - oddi + 1 <= p.rows_valid);
          if (ok) {
            const long long gco = (long long)n * p.out_cols + fo + 1;
            const int cbase = grp * COUT + h * HC + oddi * Q1 + oddg * Q2;
            uint16_t* dst = p.out + (gco * p.out_rs + to + 1) * 8 + (long long)(cbase / 8) * plane_elems;
            store_chunks<Q2 / 8>(dst, plane_elems, pk);
          }
        }
      } else if constexpr (Cfg::EPI == EPI_RELU || Cfg::EPI == EPI_GELU) {
        constexpr int HC = COUT / 2;
        static_assert((Cfg::EPI != EPI_RELU && Cfg::EPI != EPI_GELU) || HC % 32 == 0, "activation epilogue: 32-column blocks");
        const long long plane_elems = p.out_ncols * p.out_rs * 8;
        const int cbase = grp * COUT + h * HC;   // first output channel of this thread (grp = blockIdx.y output-channel group)
        for (int tt = 0; tt < Cfg::TILES; ++tt, ++it) {
          const int acc = it % NACC;
          mbar_wait(&tfull[acc], (it / NACC) & 1, 5);
          tc_fence_after();
          float v[HC];
#pragma unroll
          for (int blk = 0; blk < HC / 32; ++blk) tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + acc * NG + h * HC + blk * 32, v + blk * 32);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) acc_release<Cfg::CTA2>(&tempty[acc]);
          uint32_t pk[HC / 2];
#pragma unroll
          for (int c = 0; c < HC; c += 2) {
            float a0 = v[c] + bias[cbase + c], a1 = v[c + 1] + bias[cbase + c + 1];
            if constexpr (Cfg::EPI == EPI_GELU) {   // nn.GELU() (approximate='none'): 0.5 x (1 + erf(x / sqrt 2))
              a0 = 0.5f * a0 * (1.0f + erff(a0 * 0.70710678118654752f));
              a1 = 0.5f * a1 * (1.0f + erff(a1 * 0.70710678118654752f));
            } else {
              a0 = relu_nan(a0);
              a1 = relu_nan(a1);
            }
            pk[c >> 1] = pack_act2(a0, a1);
          }
          const int tp = 1 + 8 * tt + i;
          if (colvalid && tp <= p.rows_valid) {
            uint16_t* dst = p.out + ((long long)gc * p.out_rs + tp) * 8 + (long long)(cbase / 8) * plane_elems;
            store_chunks<HC / 8>(dst, plane_elems, pk);
          }
        }
      } else if constexpr (Cfg::EPI == EPI_SHUFFLE_MSE) {
        // dec3 (k=2,s=2 transposed conv as a 1x1 GEMM with N = (a, b, 32 channels)) + final ConvTranspose2d(32,1,k2,s2) +
        // per-utterance squared error.  This thread holds, for input position (tp, fp), the d3 vectors at
        // (2(tp-1)+h, 2(fp-1)+b), b = 0, 1; each expands to a 2x2 patch of the reconstruction, together the pixels
        // t = 4(tp-1)+2h+{0,1}, f = 4(fp-1)+{0..3}: two runs of four contiguous input samples.  d3 is rounded to fp16
        // exactly as the unfused path stores it, so both paths see the same reconstruction.
        static_assert(Cfg::EPI != EPI_SHUFFLE_MSE || (COUT == 32 && NG == 128 && Cfg::MODE == MODE_1X1 && Cfg::BAR_B >= 224), "dec3 shape");
        float acc_se = 0.0f;
        const float fb = bias[160];
        for (int tt = 0; tt < Cfg::TILES; ++tt, ++it) {
          const int acc = it % NACC;
          mbar_wait(&tfull[acc], (it / NACC) & 1, 5);
          tc_fence_after();
          float v[64];
          tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + acc * NG + h * 64, v);
          tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + acc * NG + h * 64 + 32, v + 32);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) acc_release<Cfg::CTA2>(&tempty[acc]);
          const int tp = 1 + 8 * tt + i;
          if (colvalid && tp <= p.rows_valid) {
            float rec[2][4];
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              float d[32];
#pragma unroll
              for (int c = 0; c < 32; c += 2) {   // d3 rounded to fp16 like the stored layer (pack_act2), then widened again
                const uint32_t pk = pack_act2(relu_nan(v[b * 32 + c] + bias[c]), relu_nan(v[b * 32 + c + 1] + bias[c + 1]));
                const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&pk));
                d[c] = f2.x;
                d[c + 1] = f2.y;
              }
              // 4 outputs x 32 MACs as 64 packed FFMA2 (fma.rn.f32x2: two IEEE fp32 FMAs per instruction, the d3 value broadcast, the
              // weight pair (q, q+1) of channel c read as one 64-bit uniform operand): same per-output FMA order as the scalar loop,
              // half the issue slots of the part that bounds this kernel.  Weights at bias[32 + 4 c + q], q = a2 * 2 + b2.
              uint64_t r01 = pack_f32x2(fb, fb), r23 = r01;
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                const uint64_t dd = pack_f32x2(d[c], d[c]);
                r01 = fma_f32x2(dd, pack_f32x2(bias[32 + 4 * c], bias[33 + 4 * c]), r01);
                r23 = fma_f32x2(dd, pack_f32x2(bias[34 + 4 * c], bias[35 + 4 * c]), r23);
              }
              unpack_f32x2(r01, rec[0][2 * b], rec[0][2 * b + 1]);
              unpack_f32x2(r23, rec[1][2 * b], rec[1][2 * b + 1]);
            }
            const int t0 = 4 * (tp - 1) + 2 * h, f0 = 4 * (fp - 1);
            float mu[4] = {0.f, 0.f, 0.f, 0.f}, sg[4] = {1.f, 1.f, 1.f, 1.f};
            if (p.norm_mean != nullptr) {
              const float4 m4 = *reinterpret_cast<const float4*>(p.norm_mean + f0), s4 = *reinterpret_cast<const float4*>(p.norm_sd + f0);
              mu[0] = m4.x; mu[1] = m4.y; mu[2] = m4.z; mu[3] = m4.w;
              sg[0] = s4.x; sg[1] = s4.y; sg[2] = s4.z; sg[3] = s4.w;
            }
#pragma unroll
            for (int a2 = 0; a2 < 2; ++a2) {
              const float* xr = p.x + (long long)n * p.xsn + (long long)(t0 + a2) * p.xst + (long long)f0 * p.xsf;
              float xv[4];
              if (p.x_vec4) {
                const float4 q4 = __ldg(reinterpret_cast<const float4*>(xr));
                xv[0] = q4.x; xv[1] = q4.y; xv[2] = q4.z; xv[3] = q4.w;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) xv[e] = __ldg(xr + (long long)e * p.xsf);
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float xn = xv[e];
                if (p.norm_mean != nullptr) xn = (xn - mu[e]) / sg[e];
                const float df = rec[a2][e] - xn;
                acc_se = fmaf(df, df, acc_se);
              }
            }
          }
        }
        // one partial per unit: lanes -> warp (xor shuffles), 8 warps -> two alternating slot sets in the barrier page,
        // summed in warp order by warp 0 (fixed order: the score does not depend on scheduling)
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) acc_se += __shfl_xor_sync(0xffffffffu, acc_se, o);
        float* slots = reinterpret_cast<float*>(smem + Cfg::WGT_B_AL + NSTAGE * Cfg::WIN_B_AL + 160) + 8 * (unit_seq & 1);
        if (lane == 0) slots[warp] = acc_se;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (warp == 0 && lane == 0) {
          float tot = 0.0f;
#pragma unroll
          for (int k = 0; k < 8; ++k) tot += slots[k];
          p.partial[u] = tot;
        }
        ++unit_seq;
      } else if constexpr (Cfg::EPI == EPI_SHUFFLE_ROWS) {
        // Columns of a group: h*64 + a*32 + c.  The "thread set" ts = 2*grp + h owns (column offset b, 32-channel block cblk) =
        // (ts / (COUT/32), ts % (COUT/32)) and both row offsets a.  Input row j (tp = j + 1) expands to the padded output rows
        // 2tp-1 (a = 0) and 2tp (a = 1); rows are 16 bytes apart, so the 32-byte sector {2tp-2, 2tp-1} holds (a = 1 of row
        // tp-1, a = 0 of row tp).  The a = 1 half travels one lane up (lane = row within the tile), and lanes 1..7 write whole
        // sectors with one 256-bit store per channel chunk; lane 0 writes its a = 0 half, lane 7 also its a = 1 half.  The
        // first version (EPI_SHUFFLE, one (a, b) per thread) wrote 16-byte pieces 32 bytes apart: ncu showed 29 sectors per
        // store request, twice the sector writes the data needs, and the kernels ran 3-4x over their HBM time.
        // NG = 256 holds two such 128-column sub-groups (sub), i.e. twice the columns per read of the input window.
        static_assert(Cfg::EPI != EPI_SHUFFLE_ROWS || (NG % 128 == 0 && COUT % 32 == 0 && Cfg::MODE == MODE_1X1), "shuffle-rows shape");
        constexpr int CB = COUT / 32, SUB = NG / 128;
        const long long plane_elems = p.out_ncols * p.out_rs * 8;
        const float* bs[SUB];
        uint16_t* colbase[SUB];
#pragma unroll
        for (int sub = 0; sub < SUB; ++sub) {
          const int ts = 2 * (grp * SUB + sub) + h, bq = ts / CB, cblk = ts % CB;
          bs[sub] = bias + 32 * cblk;
          colbase[sub] = p.out + ((long long)n * p.out_cols + 2 * (fp - 1) + bq + 1) * p.out_rs * 8 + (long long)(4 * cblk) * plane_elems;
        }
        for (int tt = 0; tt < Cfg::TILES; ++tt, ++it) {
          const int acc = it % NACC;
          mbar_wait(&tfull[acc], (it / NACC) & 1, 5);
          tc_fence_after();
          const int tp = 1 + 8 * tt + i;
#pragma unroll
          for (int sub = 0; sub < SUB; ++sub) {
            float v[64];
            tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + acc * NG + sub * 128 + h * 64, v);
            tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + acc * NG + sub * 128 + h * 64 + 32, v + 32);
            tmem_ld_wait();
            if (sub == SUB - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) acc_release<Cfg::CTA2>(&tempty[acc]);
            }
            uint32_t p0[16], up[16];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              p0[c >> 1] = pack_act2(relu_nan(v[c] + bs[sub][c]), relu_nan(v[c + 1] + bs[sub][c + 1]));
              up[c >> 1] = pack_act2(relu_nan(v[32 + c] + bs[sub][c]), relu_nan(v[33 + c] + bs[sub][c + 1]));
            }
            uint16_t* cb = colbase[sub];
            if (i == 7 && colvalid && tp <= p.rows_valid) {   // the a = 1 half of the tile's last row: the next tile owns the other half
#pragma unroll
              for (int k = 0; k < 4; ++k) st_global_v4(cb + (long long)k * plane_elems + (2 * tp) * 8, up[4 * k], up[4 * k + 1], up[4 * k + 2], up[4 * k + 3]);
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) up[k] = __shfl_up_sync(0xffffffffu, up[k], 1);   // lane i now holds a = 1 of row tp-1 (i > 0)
            if (colvalid) {
              if (i > 0 && tp <= p.rows_valid) {
#pragma unroll
                for (int k = 0; k < 4; ++k) st_global_v8(cb + (long long)k * plane_elems + (2 * tp - 2) * 8, up + 4 * k, p0 + 4 * k);
              } else if (i > 0 && tp - 1 <= p.rows_valid) {   // tp-1 is the last valid row
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  st_global_v4(cb + (long long)k * plane_elems + (2 * tp - 2) * 8, up[4 * k], up[4 * k + 1], up[4 * k + 2], up[4 * k + 3]);
              } else if (i == 0 && tp <= p.rows_valid) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  st_global_v4(cb + (long long)k * plane_elems + (2 * tp - 1) * 8, p0[4 * k], p0[4 * k + 1], p0[4 * k + 2], p0[4 * k + 3]);
              }
            }
          }
        }
      } else {
        // EPI_SHUFFLE: columns = (sub-quadrant, COUT); group index and column block give the 2x2 output offset (a, b):
        //   quadrant id = grp * (NG / COUT) + column block, a = qid >> 1, b = qid & 1.
        // Each thread handles the columns [h*NG/2, (h+1)*NG/2) of its row.
        constexpr int HN = NG / 2;
        constexpr int QPG = NG / COUT;              // quadrants per group
        const long long plane_elems = p.out_ncols * p.out_rs * 8;
        for (int tt = 0; tt < Cfg::TILES; ++tt, ++it) {
          const int acc = it % NACC;
          mbar_wait(&tfull[acc], (it / NACC) & 1, 5);
          tc_fence_after();
          float v[HN];
#pragma unroll
          for (int blk = 0; blk < HN / 32; ++blk) tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + acc * NG + h * HN + blk * 32, v + blk * 32);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) acc_release<Cfg::CTA2>(&tempty[acc]);
          const int tp = 1 + 8 * tt + i;
          const bool ok = colvalid && (tp <= p.rows_valid);
          constexpr int CPT = (COUT < HN) ? COUT : HN;   // channels of one quadrant held by this thread
#pragma unroll
          for (int s = 0; s < HN / CPT; ++s) {
            const int col0 = h * HN + s * CPT;            // first column of this block within the group
            const int qid = grp * QPG + col0 / COUT;
            const int c0 = col0 % COUT;                   // first output channel of the block
            uint32_t pk[CPT / 2];
#pragma unroll
            for (int c = 0; c < CPT; c += 2)
              pk[c >> 1] = pack_act2(relu_nan(v[s * CPT + c] + bias[c0 + c]), relu_nan(v[s * CPT + c + 1] + bias[c0 + c + 1]));
            if (ok) {
              const int to = 2 * (tp - 1) + (qid >> 1), fo = 2 * (fp - 1) + (qid & 1);
              const long long gco = (long long)n * p.out_cols + fo + 1;
              uint16_t* dst = p.out + (gco * p.out_rs + to + 1) * 8 + (long long)(c0 / 8) * plane_elems;
              store_chunks<CPT / 8>(dst, plane_elems, pk);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  if constexpr (Cfg::CTA2) cluster_sync_all();   // neither CTA may exit (or free TMEM) while the other can still signal it
  else __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    if constexpr (Cfg::CTA2) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <class Cfg>
static int launch_conv_tc(const CUtensorMap& tmap, const ConvParams& p, int groups, int num_sms, cudaStream_t stream) {
  static bool configured[32] = {false};
  if (dfs_first_use_on_device(configured))
    DFS_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_B));
  if (p.n_units <= 0) return DFS_OK;
  int gx = (num_sms * Cfg::OCC) / groups;
  if (gx < 1) gx = 1;
  if (gx > p.n_units) gx = p.n_units;
  if constexpr (Cfg::CTA2) {
    // clusters of 2 along x; the grid is sized from the number of pairs the device can keep resident at once
    static int max_pairs[32] = {0};
    int dev = 0;
    DFS_CUDA_CHECK(cudaGetDevice(&dev));
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_B;
    cfg.stream = stream;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (max_pairs[dev & 31] == 0) {
      cfg.gridDim = dim3(2 * num_sms, 1);
      int n = 0;
      DFS_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&n, conv_tc_kernel<Cfg>, &cfg));
      DFS_REQUIRE(n >= 1, DFS_ERR_CUDA, "no CTA pair of this kernel fits the device");
      max_pairs[dev & 31] = n;
    }
    int pairs = max_pairs[dev & 31] / groups;
    if (pairs < 1) pairs = 1;
    if (pairs > (p.n_units + 1) / 2) pairs = (p.n_units + 1) / 2;
    cfg.gridDim = dim3(2 * pairs, groups);
    DFS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<Cfg>, tmap, p));
    dfs_count_launch();
    return DFS_OK;
  }
  conv_tc_kernel<Cfg><<<dim3(gx, groups), Cfg::THREADS, Cfg::SMEM_B, stream>>>(tmap, p);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

}  // namespace dfs
