// conv1_tc.cu -- CNN2D block 1 on the tensor cores:
//   nn.Conv2d(1,32,3,p=1) + BatchNorm2d + ReLU + AvgPool2d((2,1))      /root/reference/src/model.py:15-18
//
// Cin = 1 gives K = 9 per output, not a tensor-core shape as an im2col GEMM.  Instead the conv is
// written as a *Toeplitz-in-time* GEMM whose A operand is the raw (fp16, time-major) input itself:
//   row  R = (feature column f', time block tb)          -> 8 conv outputs t = 8tb .. 8tb+7
//   col  n = jj*32 + c  (jj = time offset in the block, c = output channel), N = 256
//   K    = 3 feature taps kw  x  16 consecutive input samples x[8tb-1 .. 8tb+14][f'+kw-1]
//   B_kw[n][o] = 0.5*w'[c][o-jj][kw] if 0 <= o-jj <= 2 else 0      (BN folded, 0.5 = the (2,1) average)
// With the input stored as xT[column][1+t] (fp16, 328 samples = 41 blocks of 16 B per column) row
// R's two 16-byte K chunks are xT rows R and R+1, i.e. the SWIZZLE_NONE K-major descriptor has
// LBO = 16 B, SBO = 128 B, and the feature tap kw is a +-41-row shift of the start address.  One
// 1-D bulk copy of 212 rows x 16 B (3.4 KB) feeds a 128-row tile: 3 tcgen05.mma (M=128, N=256, K=16)
// (x 2: every weight enters as fp16 value + fp16 residual, see api.cu) produce 1024 conv outputs x 32 channels.  Epilogue: +bias, ReLU, add the two time steps of a pool
// window (columns n and n+32 of the same thread -- no shuffles), fp16, FT8 stores.
// Issued MACs are 3.5x the useful ones (Toeplitz zeros) but run ~30x faster than CUDA-core FMAs.
//
// conv1_prep_kernel converts the caller's strided fp32 features into xT.
#include "common.cuh"
#include "kernels.h"
#include "layout.cuh"
#include "xt_prep.cuh"

namespace dfs {

constexpr int kXtBlocks = 41;            // 16-byte blocks (8 samples) per xT column: samples t = -1 .. 326
constexpr int kXtLead = 48;              // zero rows before column 0 (halo of the first tile)
constexpr int kC1WinRows = 128 + 2 * kXtBlocks + 2;   // 212
constexpr int kC1WinB = kC1WinRows * 16;               // 3392
constexpr int kC1WinBAl = 3456;
constexpr int kC1Stages = 4;
constexpr int kC1WgtImgB = 3 * 256 * 16 * 2;           // 24576: one image [kw][chunk 2][256][8]
constexpr int kC1BiasOff = 2 * kC1WgtImgB;             // bias as a B operand [chunk 2][256][8]: K slot 0 = fp16(bias), slot 1 = its residual
constexpr int kC1OnesOff = kC1BiasOff + 8192;          // the matching A operand [chunk 2][128][8]: every row = (1, 1, 0, ...)
constexpr int kC1WgtB = kC1OnesOff + 4096;             // value image + residual image + bias image + ones tile
constexpr int kC1EpiWarps = 16;
constexpr int kC1Threads = (kC1EpiWarps + 3) * 32;     // 608
constexpr int kC1StageWarpB = 2 * 128 * 16;              // epilogue staging per warp: 2 planes x 128 chunks x 16 B
// SPLIT ("split" precision): a stage holds the value window and the residual window, the epilogue stages value and residual chunks
template <bool SPLIT>
struct C1Geo {
  static constexpr int kStageStride = (SPLIT ? 2 : 1) * kC1WinBAl;
  static constexpr int kWarpStageB = (SPLIT ? 2 : 1) * kC1StageWarpB;
  static constexpr int kBarOff = kC1WgtB + kC1Stages * kStageStride;
  static constexpr int kStageOff = kBarOff + 256;
  static constexpr int kSmemB = kStageOff + kC1EpiWarps * kWarpStageB + kC1EpiWarps * 32 * 4;
  static_assert(kSmemB <= 227 * 1024, "shared memory budget");
};

int64_t conv1_xt_rows(int64_t n_utts) { return kXtLead + n_utts * kCols * kXtBlocks + 128 + 2 * kXtBlocks + 16; }

// ------------------------------------------------------------------------------------------
// prep: fp32 strided features -> xT fp16 (row = 8 consecutive samples of one feature column)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv1_prep_kernel(const float* __restrict__ x, long long sn, long long st, long long sf, long long total,
                                                          uint16_t* __restrict__ xt, uint16_t* __restrict__ xt_lo) {
  // item = (n, f, blk); thread mapping chosen so that global reads coalesce for the given storage order
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int f, blk;
  long long n;
  if (sf == 1) {  // feature-contiguous storage: consecutive threads -> consecutive features
    f = (int)(idx % kF);
    blk = (int)((idx / kF) % kXtBlocks);
    n = idx / ((long long)kF * kXtBlocks);
  } else {        // time-contiguous storage (the reference's transposed view): consecutive threads -> consecutive blocks
    blk = (int)(idx % kXtBlocks);
    f = (int)((idx / kXtBlocks) % kF);
    n = idx / ((long long)kF * kXtBlocks);
  }
  const float* src = x + n * sn + (long long)f * sf;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int t = 8 * blk - 1 + e;
    v[e] = (t >= 0 && t < kT) ? src[(long long)t * st] : 0.0f;
  }
  uint16_t* dst = xt + ((long long)kXtLead + (n * kCols + f + 1) * kXtBlocks + blk) * 8;
  const uint32_t h0 = pack_act2(v[0], v[1]), h1 = pack_act2(v[2], v[3]), h2 = pack_act2(v[4], v[5]), h3 = pack_act2(v[6], v[7]);   // saturating
  st_global_v4(dst, h0, h1, h2, h3);
  if (xt_lo != nullptr)   // split precision: the rounding residuals
    st_global_v4(xt_lo + (dst - xt), pack_act2_residual(v[0], v[1], h0), pack_act2_residual(v[2], v[3], h1), pack_act2_residual(v[4], v[5], h2),
                 pack_act2_residual(v[6], v[7], h3));
}

// ------------------------------------------------------------------------------------------
// GEMM kernel
// ------------------------------------------------------------------------------------------
struct Conv1TcParams {
  const uint16_t* xt;      // xT rows (16 B each)
  const uint16_t* xt_lo;   // split precision: the fp16 rounding residuals of xt, same geometry
  const uint16_t* wpack;   // [hi | lo][kw][chunk 2][n 256][8] fp16 Toeplitz weights (value, rounding residual) | bias image | ones tile
  int n_tiles;
  int n_utts;
  uint16_t* out;           // act1, FT8P; split precision: 8 residual planes after the 8 value planes
  long long out_ncols;
  float inv_scale;         // split precision: 1 / (power-of-two scale of the weight and bias images)
};

template <bool SPLIT>
__global__ void __launch_bounds__(kC1Threads, 1) conv1_tc_kernel(const __grid_constant__ Conv1TcParams p) {
  using Geo = C1Geo<SPLIT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* wsm = smem;
  uint8_t* win0 = smem + kC1WgtB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Geo::kBarOff);
  uint8_t* stage0 = smem + Geo::kStageOff;
  uint64_t* full = bars;                 // [stages]
  uint64_t* empty = bars + kC1Stages;    // [stages]
  // An accumulator (256 TMEM columns = 8 time offsets x 32 channels) is produced and drained as two N = 128 HALVES with their own
  // barriers: the MMAs of the next tile's first half start as soon as the epilogue has read this tile's first half, i.e. while it
  // still works on the second one.  With whole-tile hand-over the 8 epilogue warps of a tile parity sat idle for the ~900 cycles of
  // their tile's MMAs + two barrier round trips every time (ncu: 21 % of the stall samples on the accumulator-ready wait).  An
  // N = 128 MMA costs half an N = 256 one (64 vs 128 cycles), so the tensor-pipe time is unchanged.
  uint64_t* tfull = empty + kC1Stages;   // [2 accumulators][2 halves]
  uint64_t* tempty = tfull + 4;          // [2][2]
  uint64_t* wbar = tempty + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kC1EpiWarps && lane == 0) {
    for (int i = 0; i < kC1Stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == kC1EpiWarps + 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kC1EpiWarps) {
    // ===================== producer =====================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(wbar, kC1WgtB);
      for (int off = 0; off < kC1WgtB; off += 4096) bulk_g2s(wsm + off, reinterpret_cast<const uint8_t*>(p.wpack) + off, 4096, wbar);
      uint32_t ws = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++ws) {
        const int stage = ws % kC1Stages;
        mbar_wait(&empty[stage], ((ws / kC1Stages) & 1) ^ 1, 21);
        mbar_arrive_expect_tx(&full[stage], (SPLIT ? 2 : 1) * kC1WinB);
        // window = xT rows [R0 - 41, R0 + 128 + 41 + 2), R0 = 128*tile, shifted by the lead margin
        const long long off = ((long long)kXtLead + 128ll * tile - kXtBlocks) * 16;
        bulk_g2s(win0 + stage * Geo::kStageStride, reinterpret_cast<const uint8_t*>(p.xt) + off, kC1WinB, &full[stage]);
        if constexpr (SPLIT) bulk_g2s(win0 + stage * Geo::kStageStride + kC1WinBAl, reinterpret_cast<const uint8_t*>(p.xt_lo) + off, kC1WinB, &full[stage]);
      }
    }
  } else if (warp == kC1EpiWarps + 1) {
    // ===================== MMA issuer =====================
    if (elect_one_sync()) {   // not `lane == 0`: see conv_tc.cuh
      constexpr uint32_t idesc = umma_idesc_f16(128, 128);
      const uint64_t b_desc0 = umma_smem_desc(smem_u32(wsm), 256 * 16, 128);
      const uint32_t b_lo0 = (uint32_t)b_desc0, b_hi = (uint32_t)(b_desc0 >> 32);
      const uint64_t a_desc0 = umma_smem_desc(smem_u32(win0), 16, 128);   // LBO = 16 B: K chunk 1 of row R is row R+1
      const uint32_t a_lo0 = (uint32_t)a_desc0, a_hi = (uint32_t)(a_desc0 >> 32);
      // the bias enters through the tensor core: ones[128 x 16] * biasB[256 x 16]^T initialises the accumulator (2 of the 16 K slots used:
      // fp16 value + residual), which takes 2 FADD per channel pair out of an epilogue that is the kernel's bottleneck
      const uint32_t ones_lo = (uint32_t)umma_smem_desc(smem_u32(wsm + kC1OnesOff), 128 * 16, 128);
      mbar_wait(wbar, 0, 22);
      uint32_t ws = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++ws) {
        const int stage = ws % kC1Stages, acc = ws & 1;
        mbar_wait(&full[stage], (ws / kC1Stages) & 1, 23);
        const uint32_t a_lo = a_lo0 + (uint32_t)(stage * (Geo::kStageStride >> 4));
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {         // half hh = time offsets 4 hh .. 4 hh + 3 = rows [128 hh, 128 hh + 128) of every B image
          mbar_wait(&tempty[2 * acc + hh], ((ws >> 1) & 1) ^ 1, 24);
          tc_fence_after();
          const uint32_t d = tmem_base + acc * 256 + hh * 128;
          const uint32_t b_lo = b_lo0 + (uint32_t)(hh * 128);   // 128 rows of 16 B
          umma_f16_lohi(d, ones_lo, a_hi, b_lo + (uint32_t)(kC1BiasOff >> 4), b_hi, idesc, 0u);
#pragma unroll
          for (int part = 0; part < 2; ++part) {   // weight value, then weight residual
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
              umma_f16_lohi(d, a_lo + (uint32_t)(kw * kXtBlocks), a_hi, b_lo + (uint32_t)((part * kC1WgtImgB + kw * 8192) >> 4), b_hi, idesc, 1u);
          }
          if constexpr (SPLIT) {   // the input's rounding residuals against the weight values
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
              umma_f16_lohi(d, a_lo + (uint32_t)(kC1WinBAl >> 4) + (uint32_t)(kw * kXtBlocks), a_hi, b_lo + (uint32_t)((kw * 8192) >> 4), b_hi, idesc, 1u);
          }
          umma_commit(&tfull[2 * acc + hh]);
        }
        umma_commit(&empty[stage]);
      }
    }
  } else if (warp < kC1EpiWarps) {
    // ===================== epilogue =====================
    const int q = warp & 3;          // TMEM lane quarter
    const int h = (warp >> 2) & 1;   // channel half: channels 16h .. 16h+15 = output planes 2h, 2h+1
    const int grp = warp >> 3;       // accumulator / tile parity this warp serves
    const long long plane_elems = p.out_ncols * kAct1RS * 8;
    // Output layout FT8P (layout.cuh): pooled time step j = 4tb + k goes to parity plane j&1, row j/2 + 1.
    // Per-warp staging: a thread owns 2 channel chunks x 2 parities x 2 rows = 8 chunks of 16 B, its two rows of a
    // (chunk, parity) plane are contiguous but 32 B apart from the next lane's, which would make every global
    // store instruction touch 32 half-used sectors.  The chunks go through shared memory (XOR-swizzled,
    // conflict-free both ways) and are stored so that lane l of store j writes chunk 32j + l of the warp's
    // contiguous 1 KB run per plane.
    uint4* stage = reinterpret_cast<uint4*>(stage0 + warp * Geo::kWarpStageB);            // [value | residual][2 chunks][2 parities][64]
    int* dsttab = reinterpret_cast<int*>(stage0 + kC1EpiWarps * Geo::kWarpStageB) + warp * 32;  // first output row of each lane
    uint32_t ws = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++ws) {
      if ((int)(ws & 1) != grp) continue;
      const int acc = grp;
      // 32-bit index arithmetic (a pass is < 2^31 / (182 * 82) utterances, checked by the launcher): the 64-bit divisions by 41 and 182
      // were ~50 instructions per tile and thread
      const uint32_t R = 128u * (uint32_t)tile + 32u * q + lane;
      const uint32_t gcx = R / (uint32_t)kXtBlocks;
      const int tb = (int)(R - gcx * kXtBlocks);
      const uint32_t n = gcx / (uint32_t)kCols;
      const int fp = (int)(gcx - n * kCols);
      const bool valid = (tb < 40) && (fp >= 1) && (fp <= kF) && ((int)n < p.n_utts);
      dsttab[lane] = valid ? (int)(gcx * kAct1RS + 2 * tb + 1) : -1;
      mbar_wait(&tfull[2 * acc], (ws >> 1) & 1, 25);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + acc * 256 + 16 * h;
      const int sw = (lane >> 2) & 1;
      // TMEM reads are software-pipelined: the loads of step k+1 are in flight while step k is converted (tcgen05.wait::ld waits for
      // ALL outstanding loads of the thread, so the next pair is issued right after each wait)
      float a[2][16], b[2][16];
      tmem_ld_32x16(taddr, a[0]);
      tmem_ld_32x16(taddr + 32, b[0]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // pooled row within the block: conv time offsets jj = 2k, 2k+1
        tmem_ld_wait();
        if (k & 1) {  // all TMEM reads of this warp from half k >> 1 are done: hand it back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[2 * acc + (k >> 1)]);
        }
        if (k == 1) {
          mbar_wait(&tfull[2 * acc + 1], (ws >> 1) & 1, 26);
          tc_fence_after();
        }
        if (k < 3) {
          tmem_ld_32x16(taddr + (2 * k + 2) * 32, a[(k + 1) & 1]);
          tmem_ld_32x16(taddr + (2 * k + 3) * 32, b[(k + 1) & 1]);
        }
        const float* av = a[k & 1];
        const float* bv = b[k & 1];
        uint32_t pk[8];
        [[maybe_unused]] uint32_t pr[8];
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          const float o0 = relu_nan(av[c]) + relu_nan(bv[c]);                 // the bias is already in the accumulator
          const float o1 = relu_nan(av[c + 1]) + relu_nan(bv[c + 1]);
          if constexpr (SPLIT) {
            const float s0 = o0 * p.inv_scale, s1 = o1 * p.inv_scale;   // exact: a power of two
            pk[c >> 1] = pack_act2(s0, s1);
            pr[c >> 1] = pack_act2_residual(s0, s1, pk[c >> 1]);
          } else {
            pk[c >> 1] = pack_act2(o0, o1);
          }
        }
        // pooled step j = 4tb + k: parity k&1, row offset k>>1 within the lane's two rows of that parity plane
        const int slot = (k & 1) * 64 + 2 * lane + ((k >> 1) ^ sw);
        stage[slot] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        stage[128 + slot] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        if constexpr (SPLIT) {
          stage[256 + slot] = make_uint4(pr[0], pr[1], pr[2], pr[3]);
          stage[384 + slot] = make_uint4(pr[4], pr[5], pr[6], pr[7]);
        }
      }
      __syncwarp();
#pragma unroll
      for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {   // value planes 0..7, residual planes 8..15
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
          for (int par = 0; par < 2; ++par) {
            uint16_t* pbase = p.out + (long long)(part * 8 + par * 4 + 2 * h + pl) * plane_elems;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int m = 32 * j + lane, r = m >> 1, c = m & 1;
              const int row = dsttab[r];
              const uint4 v = stage[part * 256 + pl * 128 + par * 64 + 2 * r + (c ^ ((r >> 2) & 1))];
              if (row >= 0) st_global_v4(pbase + (long long)(row + c) * 8, v.x, v.y, v.z, v.w);
            }
          }
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kC1EpiWarps + 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// fp32 strided features -> xT (the A operand image of conv1_tc_kernel)
static int launch_conv1_prep_any(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, uint16_t* xt, uint16_t* xt_lo, cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  if (sf == 1) {   // feature-contiguous storage: transpose through shared memory (xt_prep.cuh)
    if (xt_lo != nullptr)
      xt_prep_transpose_kernel<true><<<dim3((kF + 31) / 32, n_utts), 256, 0, stream>>>(x, sn, st, kCols, 1, kXtLead, nullptr, nullptr, xt, xt_lo);
    else
      xt_prep_transpose_kernel<false><<<dim3((kF + 31) / 32, n_utts), 256, 0, stream>>>(x, sn, st, kCols, 1, kXtLead, nullptr, nullptr, xt);
  } else {
    const long long total = (long long)n_utts * kF * kXtBlocks;
    conv1_prep_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, stream>>>(x, sn, st, sf, total, xt, xt_lo);
  }
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}
int launch_conv1_prep(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, uint16_t* xt, cudaStream_t stream) {
  return launch_conv1_prep_any(x, sn, st, sf, n_utts, xt, nullptr, stream);
}

int launch_conv1_tc(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, uint16_t* xt, const uint16_t* wpack, const float* bias_half,
                    ActBuf out, int num_sms, cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  DFS_REQUIRE(n_utts <= 100000, DFS_ERR_INVALID, "conv1: at most 100,000 utterances per pass (32-bit row indices)");
  DFS_PROPAGATE(launch_conv1_prep(x, sn, st, sf, n_utts, xt, stream));
  static bool configured[32] = {false};
  if (dfs_first_use_on_device(configured))
    DFS_CUDA_CHECK(cudaFuncSetAttribute(conv1_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1Geo<false>::kSmemB));
  Conv1TcParams p{};
  p.xt = xt;
  p.wpack = wpack;
  (void)bias_half;   // rides in wpack's bias image
  p.n_tiles = (int)ceil_div64((long long)n_utts * kCols * kXtBlocks, 128);
  p.n_utts = n_utts;
  p.out = out.ptr;
  p.out_ncols = out.ncols;
  const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
  conv1_tc_kernel<false><<<grid, kC1Threads, C1Geo<false>::kSmemB, stream>>>(p);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

// "split" precision: input, weights and output as fp16 value + fp16 residual, 10 MMAs per tile (bias, x_hi W_hi, x_hi W_lo, x_lo W_hi)
int launch_conv1_tc_split(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, uint16_t* xt, uint16_t* xt_lo, const uint16_t* wpack,
                          float inv_scale, ActBuf out, int num_sms, cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  DFS_REQUIRE(xt_lo != nullptr && out.planes == 16, DFS_ERR_INVALID, "conv1 split: residual buffers missing");
  DFS_REQUIRE(n_utts <= 100000, DFS_ERR_INVALID, "conv1: at most 100,000 utterances per pass (32-bit row indices)");
  DFS_PROPAGATE(launch_conv1_prep_any(x, sn, st, sf, n_utts, xt, xt_lo, stream));
  static bool configured[32] = {false};
  if (dfs_first_use_on_device(configured))
    DFS_CUDA_CHECK(cudaFuncSetAttribute(conv1_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1Geo<true>::kSmemB));
  Conv1TcParams p{};
  p.xt = xt;
  p.xt_lo = xt_lo;
  p.wpack = wpack;
  p.inv_scale = inv_scale;
  p.n_tiles = (int)ceil_div64((long long)n_utts * kCols * kXtBlocks, 128);
  p.n_utts = n_utts;
  p.out = out.ptr;
  p.out_ncols = out.ncols;
  const int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
  conv1_tc_kernel<true><<<grid, kC1Threads, C1Geo<true>::kSmemB, stream>>>(p);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

}  // namespace dfs
