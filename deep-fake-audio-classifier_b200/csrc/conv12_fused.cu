// conv12_fused.cu -- CNN2D blocks 1 and 2 in ONE kernel: the layer-1 activations never reach HBM.
//   Conv2d(1,32,3,p1)+BN+ReLU+AvgPool(2,1)  ->  Conv2d(32,64,3,p1)+BN+ReLU+AvgPool(2,1)      /root/reference/src/model.py:15-25
//
// Why: conv1_tc.cu writes 1.84 MB of act1 per utterance (765 MB per 416-utterance pass, which its epilogue warps get out at about half of
// what the write path takes: measured with its stores switched off the kernel needs half the time) and conv2 reads those bytes back.  Here
// conv1 is a producer stage of conv2's pipeline and the kernel is bound by its MMAs / the shared-memory port:
//
//   unit    = 14 output feature columns of one utterance (182 padded columns = 13 units: a unit never straddles utterances).
//             Its act1 window is 16 columns wide and therefore comes from ONE conv1 tile of 16 columns x 8 time blocks
//             (the SBO tiling of cae_enc1_tc.cu); conv2's M tile still has 16 column groups, groups 0 and 15 are discarded.
//   group g = time blocks 8g..8g+7 of the unit (g = 0..4), i.e. act1 pair-rows 16g+1 .. 16g+16: two passes (16 output channels
//             each) of 6 Toeplitz MMAs with N = 128 (every weight as fp16 value + residual, like conv1_tc.cu), so that conv1's
//             accumulator needs only 128 TMEM columns.
//   window w= conv2 output pair-rows 16w .. 16w+15, needs act1 pair-rows 16w .. 16w+17 = group w plus the last pair-row of
//             group w-1 (carried in a 2 KB side buffer) and the first pair-row of group w+1: no act1 value is computed twice.
//
//   warp 16     producer : one bulk copy of the unit's xT window (18 columns x 41 rows x 16 B), double buffered
//   warp 17     issuer   : conv1 group G+2 pass 0, conv2 window G (2 tiles x 24 MMAs issued interleaved; the taps of input time
//                          steps r = 1, 2 as N = 128, those of r = 0 / r = 3 as N = 64 into the half of the accumulator they reach,
//                          conv_tc.cuh HALF_TAPS), conv1 pass 1 in the middle of the window
//   warps 8-15  conv1 epilogue: TMEM -> bias/ReLU/time-pool/fp16 in registers (BEFORE the target window buffer is free), then
//                          16-byte st.shared into the K-major act1 window [8 planes][16 cols][18 rows], fence.proxy.async,
//                          mbarrier arrive.  Window G becomes ready when group G+1 has delivered its first pair-row.
//   warps 0-7   conv2 epilogue: bias/ReLU/time-pool -> act2 (FT8, the layout conv3 reads), as conv_tc.cuh EPI_PAIR_POOL
//   TMEM        conv2 accumulators 3 x 128 columns used in rotation, conv1 accumulator 128 columns
// Per window: 24 x 64 + 24 x 48 + 12 x 64 = 3,456 cycles of MMA.  Operands and epilogue formulas are the unfused path's, except that
// conv1's bias is added in the epilogue (no room for the bias / ones operand images): act2 agrees with conv1_tc + conv_tc<PAIR> to one
// fp16 ulp on a few elements per million, scores to ~1e-6 relative.
#include "common.cuh"
#include "kernels.h"
#include "layout.cuh"

namespace dfs {

constexpr int kFuCols = 16;                                // act1 window columns = conv1 tile columns
constexpr int kFuUnitsPerUtt = 13;                         // 182 / 14
constexpr int kFuRows = 18;                                // act1 window pair-rows
constexpr int kFuColB = kFuRows * 16;                      // 288
constexpr int kFuPlaneB = kFuCols * kFuColB;               // 4608
constexpr int kFuWinB = 8 * kFuPlaneB;                     // 36,864
constexpr int kFuXBlocks = 41;
constexpr int kFuXLead = 48;                               // conv1_tc.cu kXtLead
constexpr int kFuXCols = 18;
constexpr int kFuXB = kFuXCols * kFuXBlocks * 16;          // 11,808
constexpr int kFuXBAl = 11904;
constexpr int kFuW2FullB = 6 * 32 * 128 * 2;               // 49,152: taps of r = 1, 2 [tap 6][ci/8 4][n 128][8]
constexpr int kFuW2HalfB = 3 * 32 * 64 * 2;                // 12,288: taps of r = 0 (outputs dt = 0) resp. r = 3 (dt = 1) [kw 3][ci/8 4][co 64][8]
constexpr int kFuW2B = kFuW2FullB + 2 * kFuW2HalfB;        // 73,728 (the PAIR image without its zero halves)
constexpr int kFuW1B = 2 * 2 * 3 * 4096;                   // 49,152: [pass 2][value | residual][kw 3][K chunk 2][n 128][8]
constexpr int kFuOffW1 = kFuW2B;
constexpr int kFuOffX = kFuOffW1 + kFuW1B;                 // 122,880
constexpr int kFuOffA2 = kFuOffX + 2 * kFuXBAl + 512;      // 512 B guard: conv2's tap kw = 0 of column group 0 reads one column early
constexpr int kFuOffSide = kFuOffA2 + 2 * kFuWinB + 512;
constexpr int kFuSideB = 8 * kFuCols * 16;                 // 2,048
constexpr int kFuOffBar = kFuOffSide + kFuSideB;
constexpr int kFuSmemB = kFuOffBar + 256;
constexpr int kFuThreads = 18 * 32;   // 8 + 8 epilogue warps, producer (also allocates TMEM), issuer: warps are allocated in fours: 17..20 warps leave 96 registers each
static_assert(kFuSmemB <= 227 * 1024, "shared memory budget");

struct Conv12Params {
  const uint16_t* xt;      // conv1's fp16 time-major input image (conv1_tc.cu / xt_prep.cuh), 41 rows per padded column
  const uint16_t* w1pack;  // [pass][value | residual][kw][K chunk 2][n = (c / 8) * 64 + jj * 8 + c % 8][8] Toeplitz weights, 0.5 folded
  const uint16_t* w2pack;  // [r = 1, 2 taps: 6][ci/8][n 128][8] | [r = 0: kw][ci/8][co 64][8] | [r = 3: kw][ci/8][co 64][8], 0.5 folded
  float b1[32];            // 0.5 * folded conv1 bias
  float b2[64];            // 0.5 * folded conv2 bias
  int n_units;             // 13 per utterance
  int n_utts;
  uint16_t* act2;          // FT8, 8 planes
  long long act2_plane_elems;
};

__global__ void __launch_bounds__(kFuThreads, 1) conv12_fused_kernel(const __grid_constant__ Conv12Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* w2sm = smem;
  uint8_t* w1sm = smem + kFuOffW1;
  uint8_t* xsm = smem + kFuOffX;
  uint8_t* a2sm = smem + kFuOffA2;
  uint8_t* side = smem + kFuOffSide;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFuOffBar);
  uint64_t* xfull = bars;            // [2] producer -> issuer
  uint64_t* xempty = bars + 2;       // [2] issuer (commit) -> producer
  uint64_t* a1full = bars + 4;       // conv1 accumulator ready (commit)
  uint64_t* a1empty = bars + 5;      // conv1 accumulator read out (8 warps)
  uint64_t* a2full = bars + 6;       // [2] act1 window complete (8 conv1-epilogue warps)
  uint64_t* a2empty = bars + 8;      // [2] conv2 MMAs of the window retired (commit)
  uint64_t* tfull = bars + 10;       // [3] conv2 accumulator ready (commit)
  uint64_t* tempty = bars + 13;      // [3] conv2 accumulator read out (8 warps)
  uint64_t* wbar = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // the guards and both window buffers start as zeros (conv2's discarded column groups read a little outside the planes)
  for (int i = threadIdx.x; i < (2 * kFuWinB + 1024 + kFuSideB) / 16; i += kFuThreads)
    reinterpret_cast<uint4*>(a2sm - 512)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 16 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], 1);
      mbar_init(&a2full[i], 8);
      mbar_init(&a2empty[i], 1);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    mbar_init(a1full, 1);
    mbar_init(a1empty, 8);
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == 16) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc1 = tmem_base + 384;
  const int n_my_units = (p.n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // units blockIdx.x, + gridDim.x, ...
  const int n_groups = 5 * n_my_units;                                                           // = windows of this CTA

  if (warp == 16) {
    // ===================== producer: weights once, then one xT window per unit =====================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(wbar, kFuW2B + kFuW1B);
      for (int off = 0; off < kFuW2B; off += 8192) bulk_g2s(w2sm + off, reinterpret_cast<const uint8_t*>(p.w2pack) + off, 8192, wbar);
      for (int off = 0; off < kFuW1B; off += 8192) bulk_g2s(w1sm + off, reinterpret_cast<const uint8_t*>(p.w1pack) + off, 8192, wbar);
      for (int us = 0; us < n_my_units; ++us) {
        const int u = blockIdx.x + us * gridDim.x;
        const int n = u / kFuUnitsPerUtt, k = u - n * kFuUnitsPerUtt;
        const int xb = us & 1;
        mbar_wait(&xempty[xb], ((us >> 1) & 1) ^ 1, 61);
        mbar_arrive_expect_tx(&xfull[xb], kFuXB);
        // x columns f' = 14k - 1 .. 14k + 16 of utterance n (global padded column n*182 + f'), all 41 rows each
        const long long col0 = (long long)n * kCols + 14 * k - 1;
        bulk_g2s(xsm + xb * kFuXBAl, reinterpret_cast<const uint8_t*>(p.xt) + ((long long)kFuXLead + col0 * kFuXBlocks) * 16, kFuXB, &xfull[xb]);
      }
    }
  } else if (warp == 17) {
    // ===================== MMA issuer =====================
    if (elect_one_sync()) {   // not `lane == 0`: see conv_tc.cuh
      constexpr uint32_t idesc = umma_idesc_f16(128, 128);
      constexpr uint32_t idesc64 = umma_idesc_f16(128, 64);
      const uint64_t b2hd = umma_smem_desc(smem_u32(w2sm + kFuW2FullB), 64 * 16, 128);   // half-width taps: 64 rows per K chunk
      const uint32_t b2h_lo = (uint32_t)b2hd, b2h_hi = (uint32_t)(b2hd >> 32);
      const uint64_t b1d = umma_smem_desc(smem_u32(w1sm), 128 * 16, 128);
      const uint32_t b1_lo = (uint32_t)b1d, b1_hi = (uint32_t)(b1d >> 32);
      const uint64_t a1d = umma_smem_desc(smem_u32(xsm), 16, kFuXBlocks * 16);          // LBO: next time block; SBO: next column
      const uint32_t a1_lo0 = (uint32_t)a1d, a1_hi = (uint32_t)(a1d >> 32);
      const uint64_t b2d = umma_smem_desc(smem_u32(w2sm), 128 * 16, 128);
      const uint32_t b2_lo = (uint32_t)b2d, b2_hi = (uint32_t)(b2d >> 32);
      const uint64_t a2d = umma_smem_desc(smem_u32(a2sm) - kFuColB, kFuPlaneB, kFuColB);  // one column early: tap kw reads column j + kw - 1
      const uint32_t a2_lo0 = (uint32_t)a2d, a2_hi = (uint32_t)(a2d >> 32);
      mbar_wait(wbar, 0, 62);
      int issued_passes = 0;   // conv1 pass counter of this CTA: group = / 2, pass = & 1
      auto issue_conv1_pass = [&]() {
        const int PC = issued_passes, G = PC >> 1, ps = PC & 1, us = G / 5, g = G - 5 * us, xb = us & 1;
        if (g == 0 && ps == 0) mbar_wait(&xfull[xb], (us >> 1) & 1, 63);
        mbar_wait(a1empty, (PC & 1) ^ 1, 64);
        tc_fence_after();
        const uint32_t a_lo = a1_lo0 + (uint32_t)(xb * (kFuXBAl >> 4));
#pragma unroll
        for (int part = 0; part < 2; ++part)   // weight value, then weight residual
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)   // x window column 0 is f' - 1 of act1 column 0: tap kw starts kw columns in; group g at row 8g
            umma_f16_lohi(tmem_acc1, a_lo + (uint32_t)(kw * kFuXBlocks + 8 * g), a1_hi, b1_lo + (uint32_t)(((ps * 2 + part) * 3 + kw) * (4096 >> 4)),
                          b1_hi, idesc, (part | kw) != 0 ? 1u : 0u);
        umma_commit(a1full);
        if (g == 4 && ps == 1) umma_commit(&xempty[xb]);   // the unit's last conv1 MMAs have read the x window
        ++issued_passes;
      };
      // first = true : the taps of input time steps r = 1, 2 (N = 128; the very first MMA initialises the accumulator)
      // first = false: r = 0 (reaches only the pair's first output: columns [0, 64)) and r = 3 (only the second: [64, 128)) as N = 64
      auto issue_conv2_taps = [&](uint32_t a_lo_w, int acc0, int acc1, bool first) {
#pragma unroll
        for (int t = 0; t < 6; ++t) {
          const int r = first ? 1 + t / 3 : (t < 3 ? 0 : 3), kw = t % 3;
          const int par = (r == 0 || r == 2) ? 1 : 0;
          const int rowoff = (r == 0) ? 0 : (r == 3) ? 2 : 1;
          const uint32_t dcol = (!first && r == 3) ? 64u : 0u;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t b_full = b2_lo + (uint32_t)(((t * 4 + 2 * kk) * 128 * 16) >> 4);
            const uint32_t b_half = b2h_lo + (uint32_t)(((t * 4 + 2 * kk) * 64 * 16) >> 4);   // t = 0..2: the r = 0 block, 3..5: the r = 3 block behind it
#pragma unroll
            for (int m = 0; m < 2; ++m) {   // tile index innermost: consecutive MMAs alternate between the two accumulators
              const uint32_t a_off = (uint32_t)(((par * 4 + 2 * kk) * kFuPlaneB + (kw * kFuRows + rowoff + 8 * m) * 16) >> 4);
              const uint32_t d = tmem_base + (m == 0 ? acc0 : acc1) * 128 + dcol;
              if (first) umma_f16_lohi(d, a_lo_w + a_off, a2_hi, b_full, b2_hi, idesc, (t | kk) != 0 ? 1u : 0u);
              else umma_f16_lohi(d, a_lo_w + a_off, a2_hi, b_half, b2h_hi, idesc64, 1u);
            }
          }
        }
      };
      const int n_passes = 2 * n_groups;
      for (int k = 0; k < 4 && issued_passes < n_passes; ++k) issue_conv1_pass();   // groups 0 and 1
      for (int W = 0; W < n_groups; ++W) {
        if (issued_passes < n_passes) issue_conv1_pass();        // group W + 2, pass 0: runs in the gap before window W
        const int wb = W & 1;
        const int ts0 = 2 * W, ts1 = 2 * W + 1, acc0 = ts0 % 3, acc1 = ts1 % 3;
        mbar_wait(&a2full[wb], (W >> 1) & 1, 65);
        mbar_wait(&tempty[acc0], ((ts0 / 3) & 1) ^ 1, 66);
        mbar_wait(&tempty[acc1], ((ts1 / 3) & 1) ^ 1, 66);
        tc_fence_after();
        const uint32_t a_lo_w = a2_lo0 + (uint32_t)(wb * (kFuWinB >> 4));
        issue_conv2_taps(a_lo_w, acc0, acc1, true);
        if (issued_passes < n_passes && (issued_passes & 1)) issue_conv1_pass();   // group W + 2, pass 1: pass 0 has been read out by now
        issue_conv2_taps(a_lo_w, acc0, acc1, false);
        umma_commit(&tfull[acc0]);
        umma_commit(&tfull[acc1]);
        umma_commit(&a2empty[wb]);
      }
    }
  } else if (warp >= 8 && warp < 16) {
    // ===================== conv1 epilogue: TMEM -> act1 window in shared memory =====================
    const int q = warp & 3;               // TMEM lane quarter
    const int h = (warp >> 2) & 1;        // of each 16-channel pass this warp takes channels 8h .. 8h+7: chunk planes h (pass 0), 2 + h (pass 1)
    const int j = 4 * q + (lane >> 3);    // window column
    const int i = lane & 7;               // time block within the group
    // accumulator columns of a pass: n' = 64 h + 8 jj + c8 (the host's weight image orders them so): this warp's 64 columns are contiguous
    const uint32_t taddr = tmem_acc1 + ((uint32_t)(32 * q) << 16) + 64 * h;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    for (int G = 0; G <= n_groups; ++G) {
      const int us = G / 5, g = G - 5 * us;
      uint4 pk[4][2];                     // [pooled step kq][pass]
      bool real = G < n_groups;           // G == n_groups: only closes the last window (its row 17 is zero padding)
      if (real) {
        const int u = blockIdx.x + us * gridDim.x;
        const int n = u / kFuUnitsPerUtt, k = u - n * kFuUnitsPerUtt;
        const int fp = 14 * k + j;        // padded feature index of this column
        const bool colvalid = (fp >= 1) && (fp <= kF) && (n < p.n_utts);
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          mbar_wait(a1full, ps, 67);   // pass counter 2G + ps: parity = ps
          tc_fence_after();
          const int cb = 16 * ps + 8 * h;
          // two rounds of four time offsets (32 registers of accumulators in flight, the second round's loads overlap the first round's math)
          float v[2][4][8];               // [round][time offset jj & 3][channel]
          tmem_ld_32x32(taddr, &v[0][0][0]);
#pragma unroll
          for (int rd = 0; rd < 2; ++rd) {
            tmem_ld_wait();
            if (rd == 0) {
              tmem_ld_32x32(taddr + 32, &v[1][0][0]);
            } else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(a1empty);
            }
#pragma unroll
            for (int k2 = 0; k2 < 2; ++k2) {   // pooled step kq = 2 rd + k2 within the block: conv time offsets 2kq, 2kq+1
              uint32_t w[4];
#pragma unroll
              for (int c = 0; c < 8; c += 2) {
                const float o0 = relu_nan(v[rd][2 * k2][c] + p.b1[cb + c]) + relu_nan(v[rd][2 * k2 + 1][c] + p.b1[cb + c]);
                const float o1 = relu_nan(v[rd][2 * k2][c + 1] + p.b1[cb + c + 1]) + relu_nan(v[rd][2 * k2 + 1][c + 1] + p.b1[cb + c + 1]);
                w[c >> 1] = pack_act2(o0, o1);
              }
              pk[2 * rd + k2][ps] = colvalid ? make_uint4(w[0], w[1], w[2], w[3]) : zero4;   // act1's pad columns are zeros, not relu(bias)
            }
          }
        }
      }
      // 1) row 17 of the PREVIOUS window (its buffer is being assembled, nothing to wait for): this group's first pair-row
      //    (zeros across a unit boundary / at the very end).  It is the last piece window G-1 needs, so it is delivered -- and
      //    the window announced -- before this group's own rows, which have to wait for a buffer.
      uint8_t* own = a2sm + (G & 1) * kFuWinB + j * kFuColB;
      uint8_t* prev = a2sm + ((G + 1) & 1) * kFuWinB + j * kFuColB;
      if (i == 0 && G >= 1) {
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2)
#pragma unroll
          for (int par = 0; par < 2; ++par)
            *reinterpret_cast<uint4*>(prev + (par * 4 + 2 * c2 + h) * kFuPlaneB + 17 * 16) = (real && g > 0) ? pk[par][c2] : zero4;
      }
      fence_proxy_async_smem();   // every lane orders its generic-proxy stores (also the previous iteration's own rows) before the MMAs' async-proxy reads
      __syncwarp();               // one arrival per warp (256 single-thread arrivals on one mbarrier cost ~1,000 cycles)
      if (G >= 1 && lane == 0) mbar_arrive(&a2full[(G + 1) & 1]);   // window G-1 is complete
      if (!real) break;
      // 2) own rows: the buffer of window G is free once conv2's MMAs of window G-2 have retired
      if (G >= 2) mbar_wait(&a2empty[G & 1], ((G >> 1) & 1) ^ 1, 68);
      // pooled step s = 32g + 4i + kq -> FT8P parity kq & 1, pair-row 16g + 2i + (kq >> 1) + 1 = window row 2i + (kq >> 1) + 1
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int chunk = 2 * c2 + h;   // pass c2, channels 16 c2 + 8h .. + 7
#pragma unroll
        for (int par = 0; par < 2; ++par) {
          const int plane_off = (par * 4 + chunk) * kFuPlaneB;
          // a thread's two pair-rows are 32 contiguous bytes, so the 8 lanes of a column would hit every 16-byte bank group twice per
          // store (ncu: 8 wavefronts per STS.128 instead of 4); lanes 4..7 write their rows in the opposite order, which makes both
          // stores of the quarter-warp cover 8 distinct bank groups
          const int sw = (i >> 2) & 1;
          *reinterpret_cast<uint4*>(own + plane_off + (2 * i + 1 + sw) * 16) = sw ? pk[par + 2][c2] : pk[par][c2];
          *reinterpret_cast<uint4*>(own + plane_off + (2 * i + 2 - sw) * 16) = sw ? pk[par][c2] : pk[par + 2][c2];
          if (i == 0) {   // row 0 of the own window: carried last pair-row of group G-1 (zero padding for the unit's first group)
            const uint4 carry = (g > 0) ? *reinterpret_cast<const uint4*>(side + ((par * 4 + chunk) * kFuCols + j) * 16) : zero4;
            *reinterpret_cast<uint4*>(own + plane_off) = carry;
          }
        }
      }
      __syncwarp();   // the lane with i == 0 has read the carry before the lane with i == 7 of the same column replaces it
      if (i == 7) {
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2)
#pragma unroll
          for (int par = 0; par < 2; ++par)
            *reinterpret_cast<uint4*>(side + ((par * 4 + 2 * c2 + h) * kFuCols + j) * 16) = pk[par + 2][c2];
      }
    }
  } else if (warp < 8) {
    // ===================== conv2 epilogue: bias + ReLU + time pool -> act2 (FT8) =====================
    const int q = warp & 3;
    const int h = warp >> 2;              // output channels 32h .. 32h+31 = planes 4h .. 4h+3
    const int r = 32 * q + lane, j = r >> 3, i = r & 7;
    uint32_t tile_seq = 0;
    for (int W = 0; W < n_groups; ++W) {
      const int us = W / 5, w = W - 5 * us;
      const int u = blockIdx.x + us * gridDim.x;
      const int n = u / kFuUnitsPerUtt, k = u - n * kFuUnitsPerUtt;
      const int fp = 14 * k + j;
      const bool colvalid = (j >= 1) && (j <= 14) && (fp <= kF) && (n < p.n_utts);
      const long long gc = (long long)n * kCols + fp;
#pragma unroll 1
      for (int m = 0; m < 2; ++m, ++tile_seq) {
        const int acc = tile_seq % 3;
        mbar_wait(&tfull[acc], (tile_seq / 3) & 1, 69);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + acc * 128 + h * 32;
        uint32_t pk[16];
        {
          float a[32], b[32];   // conv outputs at time 2j (columns [0, 64)) and 2j + 1 (columns [64, 128)) of this thread's 32 channels
          tmem_ld_32x32(taddr, a);
          tmem_ld_32x32(taddr + 64, b);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const float o0 = relu_nan(a[c] + p.b2[h * 32 + c]) + relu_nan(b[c] + p.b2[h * 32 + c]);
            const float o1 = relu_nan(a[c + 1] + p.b2[h * 32 + c + 1]) + relu_nan(b[c + 1] + p.b2[h * 32 + c + 1]);
            pk[c >> 1] = pack_act2(o0, o1);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        if (colvalid) {
          const int row_out = 8 * (2 * w + m) + i + 1;   // pair index + 1 = padded act2 row
          uint16_t* dst = p.act2 + (gc * kAct2RS + row_out) * 8 + (long long)(4 * h) * p.act2_plane_elems;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4)
            st_global_v4(dst + c4 * p.act2_plane_elems, pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_cnn2d_conv12_fused(const uint16_t* xt, const uint16_t* w1pack, const float* b1_half, const uint16_t* w2pack, const float* b2_half, int n_utts,
                              ActBuf act2, int num_sms, cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  static bool configured[32] = {false};
  if (dfs_first_use_on_device(configured))
    DFS_CUDA_CHECK(cudaFuncSetAttribute(conv12_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemB));
  Conv12Params p{};
  p.xt = xt;
  p.w1pack = w1pack;
  p.w2pack = w2pack;
  for (int i = 0; i < 32; ++i) p.b1[i] = b1_half[i];
  for (int i = 0; i < 64; ++i) p.b2[i] = b2_half[i];
  p.n_units = n_utts * kFuUnitsPerUtt;
  p.n_utts = n_utts;
  p.act2 = act2.ptr;
  p.act2_plane_elems = act2.plane_elems();
  const int grid = p.n_units < num_sms ? p.n_units : num_sms;
  conv12_fused_kernel<<<grid, kFuThreads, kFuSmemB, stream>>>(p);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

}  // namespace dfs
