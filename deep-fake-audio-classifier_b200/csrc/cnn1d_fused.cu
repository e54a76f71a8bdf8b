// cnn1d_fused.cu -- the whole 1D-CNN scorer in ONE kernel:
//   x.transpose(1,2) -> 3 x [Conv1d(k=3,p=1) + BatchNorm1d + ReLU] -> AdaptiveAvgPool1d(1) -> Linear(128,1) [-> sigmoid]
//   /root/reference/src/model_cnn1d.py:37-46 (+ :14-35)
//
// HBM traffic = the fp32 input read (231,120 B per utterance) + 4 B of score; no activation leaves the SM.
//
// A CTA owns units of 16 utterances and walks their 41 time tiles (8 time steps each; M = 128 = 16 utterances x 8 rows).
// Per tile the three layers are chained through shared memory and TMEM:
//
//   producers (12 warps)  fp32 rows -> fp16 K-major stage [24 planes][16 utt][10 rows][8]   (as cnn1d_l1_fused.cu; the stage is
//                         ONE tile deep but split into two K halves -- planes 0..11 / 12..23 -- that are refilled
//                         independently, so the conversion of tile j+1 overlaps the MMAs of tile j in half the memory)
//   layer 1  36 MMAs  N = 32   K = 3 taps x 192   A = stage            D = acc1 (TMEM)
//   epi 1    bias + ReLU + fp16 -> window1 [4 planes][16][10 rows][8]  (rows 2..9 = this tile, rows 0..1 = carried from the last)
//   layer 2   6 MMAs  N = 64   K = 3 x 32         A = window1          D = acc2
//   epi 2    bias + ReLU + fp16 -> window2 [8 planes][16][10][8]
//   layer 3  12 MMAs  N = 128  K = 3 x 64         A = window2          D = acc3
//   epi 3    bias + ReLU, dot with the classifier weights, running sum per (utterance, row) in ONE register;
//            after the unit's last tile: reduce over the 8 rows, / 321, + fc bias [, sigmoid], store the score
//
// Each layer runs ONE ROW BEHIND the previous one, so that the 10-row window a tile needs (8 rows + the k = 3 halo) is exactly
// "the 8 rows the previous layer just produced + the last 2 rows of its previous tile":
//   layer-1 tile j = time steps 8j .. 8j+7,  layer-2 tile j = 8j-1 .. 8j+6,  layer-3 tile j = 8j-2 .. 8j+5.
// Rows outside [0, 321) are stored as zeros (the Conv1d zero padding) / left out of the sum.  The intermediate activations are
// rounded to fp16 exactly like the stored layers of cnn1d_tc.cu, so the logits agree with the unfused path up to the
// summation order of the time mean.
//
// The single MMA-issuing thread interleaves the three layers one tile apart (layer 1 of tile k, layer 2 of tile k-1, layer 3 of
// tile k-2): whatever it waits for was produced an iteration earlier.
//
// Warps: 0-3 epi 1, 4-7 epi 2, 8-11 epi 3 (TMEM lane quarter = warp % 4), 12-23 producers, 24 MMA issuer, 25 TMEM allocator.
#include "common.cuh"
#include "kernels.h"
#include "layout.cuh"

namespace dfs {

namespace {

constexpr int kFuRows = 10;                                // window rows: 8 + the k = 3 halo
constexpr int kFuPlaneB = kColTile * kFuRows * 16 + 16;    // 2576: +16 B keeps the 16-byte st.shared of the producers conflict-free
constexpr int kFuTiles = 41;                               // 328 / 8 time tiles per unit
constexpr int kFuW1B = 3 * 24 * 32 * 16;                   // 36,864  [tap][24][32][8]
constexpr int kFuW2B = 3 * 4 * 64 * 16;                    // 12,288  [tap][4][64][8]
constexpr int kFuW3B = 3 * 8 * 128 * 16;                   // 49,152  [tap][8][128][8]
constexpr int kFuStageB = 24 * kFuPlaneB;                  // 61,824
constexpr int kFuHalfB = 12 * kFuPlaneB;                   // 30,912
constexpr int kFuWin1B = 4 * kFuPlaneB;                    // 10,304
constexpr int kFuWin2B = 8 * kFuPlaneB;                    // 20,608
constexpr int kFuCarry1B = 4 * kColTile * 2 * 16;          // 2,048  [plane][utt][2 rows][16 B]
constexpr int kFuCarry2B = 8 * kColTile * 2 * 16;          // 4,096
constexpr int kFuOffW2 = kFuW1B;
constexpr int kFuOffW3 = kFuOffW2 + kFuW2B;
constexpr int kFuOffStage = kFuOffW3 + kFuW3B;             // 98,304 (1024-aligned)
constexpr int kFuOffWin1 = kFuOffStage + kFuStageB;
constexpr int kFuOffWin2 = kFuOffWin1 + kFuWin1B;
constexpr int kFuOffCarry1 = kFuOffWin2 + kFuWin2B;
constexpr int kFuOffCarry2 = kFuOffCarry1 + kFuCarry1B;
constexpr int kFuOffBar = (kFuOffCarry2 + kFuCarry2B + 127) & ~127;
constexpr int kFuSmemB = kFuOffBar + 256;
constexpr int kFuProdWarp0 = 12, kFuProdWarps = 12;
constexpr int kFuMmaWarp = kFuProdWarp0 + kFuProdWarps;    // 24
constexpr int kFuThreads = 32 * (kFuMmaWarp + 2);          // 832
constexpr int kFuAcc1 = 0, kFuAcc2 = 64, kFuAcc3 = 192;    // TMEM column offsets: 2 x 32, 2 x 64, 2 x 128 -> 448 of 512
static_assert(kFuSmemB <= 227 * 1024, "shared memory budget");
static_assert(kFuOffStage % 1024 == 0, "stage alignment");

struct FusedParams {
  const float* x;          // dense [n][321][180] fp32
  long long sn;            // utterance stride in elements
  const uint16_t* w1;      // [tap 3][24][32][8] fp16, BN folded
  const uint16_t* w2;      // [tap 3][4][64][8]
  const uint16_t* w3;      // [tap 3][8][128][8]
  float b1[32], b2[64], b3[128];
  float fcw[128];          // classifier weight
  float fcb;
  int apply_sigmoid;
  int n_units;             // 16-utterance column tiles
  int n_utts;
  float* out;              // [n] logits / scores
};

// ===================== epi 1 / epi 2: bias + ReLU + fp16 -> the next layer's window =====================
// LAYER 1: 32 channels = 4 planes, tile j = time steps 8j .. 8j+7; LAYER 2: 64 channels = 8 planes, tile j = 8j-1 .. 8j+6.
template <int LAYER>
__device__ __forceinline__ void fused_epi_relu(const FusedParams& p, uint8_t* smem, uint32_t tmem_base, int total, int qd, int lane) {
  constexpr bool second = LAYER == 2;
  constexpr int NBLK = second ? 2 : 1;                                // 32-channel blocks (4 planes each)
  constexpr uint32_t ACC_COL = second ? kFuAcc2 : kFuAcc1;
  constexpr int ACC_STRIDE = second ? 64 : 32;
  constexpr int LAG = second ? 1 : 0;
  uint8_t* win = smem + (second ? kFuOffWin2 : kFuOffWin1);
  uint8_t* carry = smem + (second ? kFuOffCarry2 : kFuOffCarry1);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFuOffBar);
  uint64_t* tfull = bars + (second ? 8 : 4);
  uint64_t* tempty = bars + (second ? 10 : 6);
  uint64_t* wfull = bars + (second ? 18 : 16);
  uint64_t* wfree = bars + (second ? 19 : 17);
  const int r = 32 * qd + lane, g = r >> 3, i = r & 7;               // TMEM lane = (utterance g of the unit, row i of the tile)
  for (int k = 0; k < total; ++k) {
    const int j = k % kFuTiles;
    const uint32_t acc = (uint32_t)(k & 1);
    const int t = 8 * j - LAG + i;
    const bool tvalid = t >= 0 && t < kT;
    mbar_wait(&tfull[acc], (k >> 1) & 1, 69);
    // the window is free once the next layer's MMAs of the previous tile have completed
    mbar_wait(wfree, (k & 1) ^ 1, 70);
    tc_fence_after();
    // rows 0..1 = the last two rows of the previous tile (zeros at the start of a unit), handed over through `carry`
    if (i < 2) {
#pragma unroll
      for (int pl = 0; pl < 4 * NBLK; ++pl) {
        uint4 cv = make_uint4(0, 0, 0, 0);
        if (j != 0) cv = *reinterpret_cast<const uint4*>(carry + ((pl * kColTile + g) * 2 + i) * 16);
        *reinterpret_cast<uint4*>(win + pl * kFuPlaneB + (g * kFuRows + i) * 16) = cv;
      }
    }
    __syncwarp();
#pragma unroll
    for (int blk = 0; blk < NBLK; ++blk) {
      float v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(32 * qd) << 16) + ACC_COL + acc * ACC_STRIDE + blk * 32, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        const float b0 = second ? p.b2[blk * 32 + c] : p.b1[c], b1 = second ? p.b2[blk * 32 + c + 1] : p.b1[c + 1];
        pk[c >> 1] = pack_act2(tvalid ? relu_nan(v[c] + b0) : 0.0f, tvalid ? relu_nan(v[c + 1] + b1) : 0.0f);
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int pl = 4 * blk + q4;
        const uint4 o = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
        *reinterpret_cast<uint4*>(win + pl * kFuPlaneB + (g * kFuRows + 2 + i) * 16) = o;
        if (i >= 6) *reinterpret_cast<uint4*>(carry + ((pl * kColTile + g) * 2 + (i - 6)) * 16) = o;
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();     // generic-proxy stores -> visible to the tensor core's async-proxy reads
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&tempty[acc]);
      mbar_arrive(wfull);
    }
  }
}

}  // namespace

__global__ void __launch_bounds__(kFuThreads, 1) cnn1d_fused_kernel(const __grid_constant__ FusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* w1s = smem;
  uint8_t* w2s = smem + kFuOffW2;
  uint8_t* w3s = smem + kFuOffW3;
  uint8_t* stage = smem + kFuOffStage;
  uint8_t* win1 = smem + kFuOffWin1;
  uint8_t* win2 = smem + kFuOffWin2;
  uint8_t* carry1 = smem + kFuOffCarry1;
  uint8_t* carry2 = smem + kFuOffCarry2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFuOffBar);
  uint64_t* fullh = bars;            // [2]  producers of K half h -> MMA (6 warp arrivals)
  uint64_t* emptyh = bars + 2;       // [2]  MMA -> producers of half h
  uint64_t* tfull1 = bars + 4;       // [2]  layer-1 accumulator ready
  uint64_t* tempty1 = bars + 6;      // [2]  ... drained (4 warps)
  uint64_t* tfull2 = bars + 8;
  uint64_t* tempty2 = bars + 10;
  uint64_t* tfull3 = bars + 12;
  uint64_t* tempty3 = bars + 14;
  uint64_t* w1full = bars + 16;      // window1 written by epi 1 (4 warps)
  uint64_t* w1free = bars + 17;      // layer-2 MMAs that read window1 have completed
  uint64_t* w2full = bars + 18;
  uint64_t* w2free = bars + 19;
  uint64_t* wbar = bars + 20;        // weights resident
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // zero the stage (plane 23, the upper half of plane 22 and the pads are never written again), the windows and the carries
  for (int i = threadIdx.x; i < (kFuOffBar - kFuOffStage) / 16; i += kFuThreads) reinterpret_cast<uint4*>(stage)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == kFuMmaWarp && lane == 0) {
    for (int h = 0; h < 2; ++h) { mbar_init(&fullh[h], kFuProdWarps / 2); mbar_init(&emptyh[h], 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull1[a], 1); mbar_init(&tempty1[a], 4);
      mbar_init(&tfull2[a], 1); mbar_init(&tempty2[a], 4);
      mbar_init(&tfull3[a], 1); mbar_init(&tempty3[a], 4);
    }
    mbar_init(w1full, 4); mbar_init(w1free, 1);
    mbar_init(w2full, 4); mbar_init(w2free, 1);
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == kFuMmaWarp + 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // tiles this CTA walks: units blockIdx.x, blockIdx.x + gridDim.x, ...; global tile counter k = (unit ordinal) * 41 + j
  const int my_units = (p.n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = my_units * kFuTiles;

  if (warp >= kFuProdWarp0 && warp < kFuMmaWarp) {
    // ===================== producers: fp32 rows -> fp16 K-major stage, one K half per 6 warps =====================
    // A 720-byte feature row is 45 pieces of 16 B (4 features).  Thread q of a half owns ONE piece index and the two utterance columns
    // cg, cg + 8, and walks their 10 rows: consecutive lanes read consecutive 16-byte pieces, so a warp's load instruction covers whole
    // contiguous sectors (the first mapping gave a lane 32 bytes as two 16-byte loads 32 B apart: every sector was fetched by two
    // instructions, L1 throughput 77 %, 30 sectors per request).  A piece converts to 8 bytes = half of a 16-byte K chunk: st.shared.v2.
    const int pt = threadIdx.x - 32 * kFuProdWarp0;      // 0 .. 383
    const int half = pt >= 192;
    const int q = half ? pt - 192 : pt;
    const int piece = q % 24, cg = q / 24;               // cg 0..7
    const bool active = half ? piece < 21 : true;        // half 1 = pieces 24 .. 44 (features 96 .. 179); plane 22's upper half and plane 23 stay zero
    const int gp = (half ? 24 : 0) + piece;              // piece index within the row
    const int c8 = gp >> 1, sub = gp & 1;
    if (pt == 0) {
      mbar_arrive_expect_tx(wbar, kFuW1B + kFuW2B + kFuW3B);
      for (int off = 0; off < kFuW1B; off += 12288) bulk_g2s(w1s + off, reinterpret_cast<const uint8_t*>(p.w1) + off, 12288, wbar);
      bulk_g2s(w2s, p.w2, kFuW2B, wbar);
      for (int off = 0; off < kFuW3B; off += 16384) bulk_g2s(w3s + off, reinterpret_cast<const uint8_t*>(p.w3) + off, 16384, wbar);
    }
    // L2 prefetch one tile ahead: thread pt < 16 asks for the 8 new rows (5,760 contiguous bytes) of its utterance
    auto prefetch_rows = [&](int u, int tt) {
      if (pt >= kColTile || u >= p.n_units) return;
      const long long gn = (long long)kColTile * u + pt;
      if (gn >= p.n_utts) return;
      const int t0 = 8 * tt, t1 = (8 * tt + 8) < kT ? (8 * tt + 8) : kT;
      if (t0 >= t1) return;
      const float* src = p.x + gn * p.sn + (long long)t0 * kF;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)((t1 - t0) * kF * 4)) : "memory");
    };
    prefetch_rows(blockIdx.x, 0);
    uint32_t k = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const long long gn0 = (long long)kColTile * u + cg, gn1 = gn0 + 8;
      const bool v0 = active && gn0 < p.n_utts, v1 = active && gn1 < p.n_utts;
      const float* base0 = p.x + (v0 ? gn0 : 0) * p.sn + 4 * gp;
      const float* base1 = p.x + (v1 ? gn1 : 0) * p.sn + 4 * gp;
      uint8_t* dst0 = stage + c8 * kFuPlaneB + cg * (kFuRows * 16) + 8 * sub;
      uint8_t* dst1 = dst0 + 8 * (kFuRows * 16);
      for (int tt = 0; tt < kFuTiles; ++tt, ++k) {
        if (tt + 1 < kFuTiles) prefetch_rows(u, tt + 1);
        else prefetch_rows(u + gridDim.x, 0);
        // The half-stage is ONE tile deep: it may be overwritten only after the MMAs of the previous tile have read it.  The global
        // loads do not touch it, so the first five rows are requested BEFORE that wait and fly while the previous tile's MMAs run.
        constexpr int U = 5;
        float4 ra[U], rb[U];
        auto load_rows = [&](int r0) {
#pragma unroll
          for (int e = 0; e < U; ++e) {
            const int t = 8 * tt - 1 + r0 + e;
            ra[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            rb[e] = ra[e];
            if (t >= 0 && t < kT) {
              if (v0) ra[e] = __ldg(reinterpret_cast<const float4*>(base0 + (long long)t * kF));
              if (v1) rb[e] = __ldg(reinterpret_cast<const float4*>(base1 + (long long)t * kF));
            }
          }
        };
        auto store_rows = [&](int r0) {
#pragma unroll
          for (int e = 0; e < U; ++e) {
            *reinterpret_cast<uint2*>(dst0 + (r0 + e) * 16) = make_uint2(pack_act2(ra[e].x, ra[e].y), pack_act2(ra[e].z, ra[e].w));
            *reinterpret_cast<uint2*>(dst1 + (r0 + e) * 16) = make_uint2(pack_act2(rb[e].x, rb[e].y), pack_act2(rb[e].z, rb[e].w));
          }
        };
        if (active) load_rows(0);
        mbar_wait(&emptyh[half], (k & 1) ^ 1, 61);
        if (active) {
          store_rows(0);
          load_rows(U);
          store_rows(U);
        }
        fence_proxy_async_smem();   // every lane: generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&fullh[half]);
      }
    }
  } else if (warp == kFuMmaWarp) {
    // ===================== MMA issuer: layer 1 of tile k, layer 2 of tile k-1, layer 3 of tile k-2 =====================
    if (elect_one_sync()) {   // not `lane == 0`: see conv_tc.cuh
      constexpr uint32_t idesc1 = umma_idesc_f16(128, 32), idesc2 = umma_idesc_f16(128, 64), idesc3 = umma_idesc_f16(128, 128);
      const uint64_t a1d = umma_smem_desc(smem_u32(stage), kFuPlaneB, kFuRows * 16);
      const uint64_t a2d = umma_smem_desc(smem_u32(win1), kFuPlaneB, kFuRows * 16);
      const uint64_t a3d = umma_smem_desc(smem_u32(win2), kFuPlaneB, kFuRows * 16);
      const uint64_t b1d = umma_smem_desc(smem_u32(w1s), 32 * 16, 128);     // LBO = K-chunk stride = rows x 16 B
      const uint64_t b2d = umma_smem_desc(smem_u32(w2s), 64 * 16, 128);
      const uint64_t b3d = umma_smem_desc(smem_u32(w3s), 128 * 16, 128);
      const uint32_t a1lo = (uint32_t)a1d, a1hi = (uint32_t)(a1d >> 32), a2lo = (uint32_t)a2d, a2hi = (uint32_t)(a2d >> 32);
      const uint32_t a3lo = (uint32_t)a3d, a3hi = (uint32_t)(a3d >> 32);
      const uint32_t b1lo = (uint32_t)b1d, b1hi = (uint32_t)(b1d >> 32), b2lo = (uint32_t)b2d, b2hi = (uint32_t)(b2d >> 32);
      const uint32_t b3lo = (uint32_t)b3d, b3hi = (uint32_t)(b3d >> 32);
      mbar_wait(wbar, 0, 62);
      for (int k = 0; k < total + 2; ++k) {
        if (k < total) {   // ---- layer 1, tile k: two K halves of 18 MMAs
          const uint32_t acc = (uint32_t)(k & 1);
          mbar_wait(&tempty1[acc], ((k >> 1) & 1) ^ 1, 63);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&fullh[h], k & 1, 64);
            tc_fence_after();
#pragma unroll
            for (int tap = 0; tap < 3; ++tap) {
#pragma unroll
              for (int kk = 6 * h; kk < 6 * h + 6; ++kk) {
                const uint32_t a_off = (uint32_t)((2 * kk * kFuPlaneB + tap * 16) >> 4);
                const uint32_t b_off = (uint32_t)(((tap * 24 + 2 * kk) * 32 * 16) >> 4);
                umma_f16_lohi(tmem_base + kFuAcc1 + acc * 32, a1lo + a_off, a1hi, b1lo + b_off, b1hi, idesc1, (h | tap | (kk - 6 * h)) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&emptyh[h]);
          }
          umma_commit(&tfull1[acc]);
        }
        if (k >= 1 && k - 1 < total) {   // ---- layer 2, tile k-1
          const int k2 = k - 1;
          const uint32_t acc = (uint32_t)(k2 & 1);
          mbar_wait(&w1full[0], k2 & 1, 65);
          mbar_wait(&tempty2[acc], ((k2 >> 1) & 1) ^ 1, 66);
          tc_fence_after();
#pragma unroll
          for (int tap = 0; tap < 3; ++tap) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint32_t a_off = (uint32_t)((2 * kk * kFuPlaneB + tap * 16) >> 4);
              const uint32_t b_off = (uint32_t)(((tap * 4 + 2 * kk) * 64 * 16) >> 4);
              umma_f16_lohi(tmem_base + kFuAcc2 + acc * 64, a2lo + a_off, a2hi, b2lo + b_off, b2hi, idesc2, (tap | kk) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&w1free[0]);
          umma_commit(&tfull2[acc]);
        }
        if (k >= 2) {   // ---- layer 3, tile k-2
          const int k3 = k - 2;
          const uint32_t acc = (uint32_t)(k3 & 1);
          mbar_wait(&w2full[0], k3 & 1, 67);
          mbar_wait(&tempty3[acc], ((k3 >> 1) & 1) ^ 1, 68);
          tc_fence_after();
#pragma unroll
          for (int tap = 0; tap < 3; ++tap) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t a_off = (uint32_t)((2 * kk * kFuPlaneB + tap * 16) >> 4);
              const uint32_t b_off = (uint32_t)(((tap * 8 + 2 * kk) * 128 * 16) >> 4);
              umma_f16_lohi(tmem_base + kFuAcc3 + acc * 128, a3lo + a_off, a3hi, b3lo + b_off, b3hi, idesc3, (tap | kk) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&w2free[0]);
          umma_commit(&tfull3[acc]);
        }
      }
    }
  } else if (warp < 4) {
    fused_epi_relu<1>(p, smem, tmem_base, total, warp & 3, lane);
  } else if (warp < 8) {
    fused_epi_relu<2>(p, smem, tmem_base, total, warp & 3, lane);
  } else if (warp < 12) {
    // ===================== epi 3: bias + ReLU, classifier dot, time mean =====================
    const int qd = warp & 3;
    const int r = 32 * qd + lane, g = r >> 3, i = r & 7;
    float run = 0.0f;
    int ord = 0;                                                      // unit ordinal of this CTA
    for (int k = 0; k < total; ++k) {
      const int j = k % kFuTiles;
      const uint32_t acc = (uint32_t)(k & 1);
      const int t = 8 * j - 2 + i;                                    // layer-3 tile j holds time steps 8j-2 .. 8j+5
      const bool tvalid = t >= 0 && t < kT;
      mbar_wait(&tfull3[acc], (k >> 1) & 1, 71);
      tc_fence_after();
      float part = 0.0f;
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(32 * qd) << 16) + kFuAcc3 + acc * 128 + blk * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) part = fmaf(relu_nan(v[c] + p.b3[blk * 32 + c]), p.fcw[blk * 32 + c], part);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty3[acc]);
      if (tvalid) run += part;
      if (j == kFuTiles - 1) {
        float s = run;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        const long long n = (long long)kColTile * ((long long)blockIdx.x + (long long)ord * gridDim.x) + g;
        if (i == 0 && n < p.n_utts) {
          const float z = s / (float)kT + p.fcb;
          p.out[n] = p.apply_sigmoid ? 1.0f / (1.0f + expf(-z)) : z;
        }
        run = 0.0f;
        ++ord;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kFuMmaWarp + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_cnn1d_fused(const float* x, int64_t sn, int n_utts, const uint16_t* w1, const uint16_t* w2, const uint16_t* w3, const float* b1,
                       const float* b2, const float* b3, const float* fcw_host, float fcb, int apply_sigmoid, float* out, int num_sms,
                       cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  static bool configured[32] = {false};
  if (dfs_first_use_on_device(configured))
    DFS_CUDA_CHECK(cudaFuncSetAttribute(cnn1d_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemB));
  FusedParams p{};
  p.x = x;
  p.sn = sn;
  p.w1 = w1;
  p.w2 = w2;
  p.w3 = w3;
  for (int i = 0; i < 32; ++i) p.b1[i] = b1[i];
  for (int i = 0; i < 64; ++i) p.b2[i] = b2[i];
  for (int i = 0; i < 128; ++i) { p.b3[i] = b3[i]; p.fcw[i] = fcw_host[i]; }
  p.fcb = fcb;
  p.apply_sigmoid = apply_sigmoid;
  p.n_units = (n_utts + kColTile - 1) / kColTile;
  p.n_utts = n_utts;
  p.out = out;
  const int grid = p.n_units < num_sms ? p.n_units : num_sms;
  cnn1d_fused_kernel<<<grid, kFuThreads, kFuSmemB, stream>>>(p);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

}  // namespace dfs
