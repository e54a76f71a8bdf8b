// cae_enc1_tc.cu -- CAE encoder block 1 on the tensor cores:
//   FeatureNormalizer.transform + nn.Conv2d(1,32,3,p=1) + BatchNorm2d + ReLU + AvgPool2d(2)
//   /root/reference/src/dataset_cae.py:37-41, /root/reference/src/model_cae.py:33-38
//
// Same Toeplitz-in-time GEMM as the 2D-CNN's first layer (conv1_tc.cu): a GEMM row = 8 consecutive conv outputs in time of
// one feature column, N = 256 = (time offset jj, output channel c), K = 16 consecutive input samples per feature tap,
// three taps -> 3 tcgen05.mma (M=128, N=256, K=16) per 1024 conv positions.  The input image is column-major,
//     xT2[(gc * 41 + tb)] = 8 samples x[8tb-1 .. 8tb+6][f]  (normalised, fp16),  gc = n * 184 + f + 2
// (pad columns 0, 1, 182, 183 of every utterance are zero), so K chunk 1 of a row is the next row (LBO = 16 B) and a
// feature tap is a +-41-row shift of the descriptor start address.  The M = 128 rows of a tile are 16 COLUMNS x 8 TIME
// BLOCKS: the stride between the 8-row core-matrix groups (SBO) is free, and SBO = 41 rows walks across columns.  That
// puts the 2x2 pool's feature partner 8 TMEM lanes away (one __shfl_xor 8, as in conv_tc's EPI_POOL_TF), the time partner
// in the same thread, and gives every 8 consecutive lanes 8 consecutive time blocks of one column = 256 contiguous bytes
// of each output plane.  One bulk copy of 18 columns x 41 rows (11.8 KB) feeds the 5 tiles (tb 0..39) of a 16-column unit.
// The 0.25 of the 2x2 average is folded into weights and bias (ReLU is positively homogeneous).  Output: e1 in the FT8P
// layout enc2's PAIR GEMM reads (layout.cuh): pooled time step 4tb + k -> parity plane k & 1, row 2tb + (k >> 1) + 1.
// Like conv1_tc the kernel is bound by its epilogue (a chain of dependent latencies), not by the MMAs; the bias is added by
// the tensor core (a ones x bias MMA initialises every accumulator).
#include "common.cuh"
#include "kernels.h"
#include "layout.cuh"
#include "xt_prep.cuh"

namespace dfs {

constexpr int kE1Cols = 184;                            // padded feature columns per utterance: f'' = f + 2
constexpr int kE1Blocks = 41;                           // 16-byte rows (8 samples) per column: samples t = -1 .. 326
constexpr int kE1Lead = 48;                             // zero rows before column 0 (tap -1 of the first unit)
constexpr int kE1UnitCols = 16;
constexpr int kE1WinRows = (kE1UnitCols + 2) * kE1Blocks;   // 738
constexpr int kE1WinB = kE1WinRows * 16;                // 11808
constexpr int kE1WinBAl = 12288;
constexpr int kE1Stages = 4;
constexpr int kE1TilesPerUnit = 5;                      // time blocks 0..39 in tiles of 8 (block 40 only feeds K chunk 1)
constexpr int kE1BiasOff = 3 * 256 * 16 * 2;            // 24576: bias as a B operand [chunk 2][256][8] (K slot 0 = fp16(bias), slot 1 = its residual)
constexpr int kE1OnesOff = kE1BiasOff + 8192;           // the matching A operand [chunk 2][128][8]: every row = (1, 1, 0, ...)
constexpr int kE1WgtB = kE1OnesOff + 4096;              // weights + bias image + ones tile
constexpr int kE1EpiWarps = 16;
constexpr int kE1Threads = (kE1EpiWarps + 3) * 32;      // 608
constexpr int kE1BarOff = kE1WgtB + kE1Stages * kE1WinBAl;
constexpr int kE1SmemB = kE1BarOff + 256;

int64_t cae_enc1_xt_rows(int64_t n_utts) { return kE1Lead + (n_utts * kE1Cols + 2 * kE1UnitCols + 2) * kE1Blocks + 16; }

// fp32 strided features -> normalised fp16 xT2 rows; pad columns / lead / tail rows are zeroed once at allocation
__global__ void __launch_bounds__(256) cae_enc1_prep_kernel(const float* __restrict__ x, long long sn, long long st, long long sf, long long total,
                                                             const float* __restrict__ mean, const float* __restrict__ sd,
                                                             uint16_t* __restrict__ xt) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int f, tb;
  long long n;
  if (sf <= st) {  // feature-contiguous storage: consecutive threads -> consecutive features (coalesced reads)
    f = (int)(idx % kF);
    tb = (int)((idx / kF) % kE1Blocks);
    n = idx / ((long long)kF * kE1Blocks);
  } else {         // time-contiguous storage (the reference's transposed view): consecutive threads -> consecutive blocks
    tb = (int)(idx % kE1Blocks);
    f = (int)((idx / kE1Blocks) % kF);
    n = idx / ((long long)kF * kE1Blocks);
  }
  const float* src = x + n * sn + (long long)f * sf;
  const float m = mean != nullptr ? mean[f] : 0.0f, s = sd != nullptr ? sd[f] : 1.0f;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int t = 8 * tb - 1 + e;
    float val = 0.0f;                               // zero padding is applied AFTER the normalisation
    if (t >= 0 && t < kT) {
      val = src[(long long)t * st];
      if (mean != nullptr) val = (val - m) / s;      // pack_act2 saturates to +-65504
    }
    v[e] = val;
  }
  uint16_t* dst = xt + ((long long)kE1Lead + (n * kE1Cols + f + 2) * kE1Blocks + tb) * 8;
  st_global_v4(dst, pack_act2(v[0], v[1]), pack_act2(v[2], v[3]), pack_act2(v[4], v[5]), pack_act2(v[6], v[7]));
}

struct Enc1TcParams {
  const uint16_t* xt;      // xT2 rows (16 B each)
  const uint16_t* wpack;   // [kw][chunk 2][n 256][8] fp16 Toeplitz weights, 0.25 folded | bias image | ones tile
  int n_units;             // 16-column units over the global column index n * 184 + f''
  int n_utts;
  uint16_t* out;           // e1, FT8P, 8 planes x (92 columns per utterance) x RS 82
  long long out_plane_elems;
  int out_rs;
  int out_cols;
};

__global__ void __launch_bounds__(kE1Threads, 1) cae_enc1_tc_kernel(const __grid_constant__ Enc1TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* wsm = smem;
  uint8_t* win0 = smem + kE1WgtB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kE1BarOff);
  uint64_t* full = bars;                 // [stages]
  uint64_t* empty = bars + kE1Stages;    // [stages]
  uint64_t* tfull = empty + kE1Stages;   // [2 accumulators][2 halves]
  uint64_t* tempty = tfull + 4;          // [2][2]
  uint64_t* wbar = tempty + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kE1EpiWarps && lane == 0) {
    for (int i = 0; i < kE1Stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }   // [2 accumulators][2 halves], see conv1_tc.cu
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == kE1EpiWarps + 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kE1EpiWarps) {
    // ===================== producer =====================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(wbar, kE1WgtB);
      for (int off = 0; off < kE1WgtB; off += 4096) bulk_g2s(wsm + off, reinterpret_cast<const uint8_t*>(p.wpack) + off, 4096, wbar);
      uint32_t ws = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++ws) {
        const int stage = ws % kE1Stages;
        mbar_wait(&empty[stage], ((ws / kE1Stages) & 1) ^ 1, 41);
        mbar_arrive_expect_tx(&full[stage], kE1WinB);
        // window = columns 16u - 1 .. 16u + 16, all 41 rows each
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.xt) + ((long long)kE1Lead + (16ll * u - 1) * kE1Blocks) * 16;
        bulk_g2s(win0 + stage * kE1WinBAl, src, kE1WinB, &full[stage]);
      }
    }
  } else if (warp == kE1EpiWarps + 1) {
    // ===================== MMA issuer =====================
    if (elect_one_sync()) {   // not `lane == 0`: see conv_tc.cuh
      constexpr uint32_t idesc = umma_idesc_f16(128, 128);   // an accumulator is produced as two N = 128 halves (time offsets 0..3 | 4..7)
      const uint64_t b_desc0 = umma_smem_desc(smem_u32(wsm), 256 * 16, 128);
      const uint32_t b_lo0 = (uint32_t)b_desc0, b_hi = (uint32_t)(b_desc0 >> 32);
      // A: K chunk 1 of a row is the next row (LBO = 16 B); the 16 core-matrix groups of a tile are 16 columns (SBO = 41 rows)
      const uint64_t a_desc0 = umma_smem_desc(smem_u32(win0), 16, kE1Blocks * 16);
      const uint32_t a_lo0 = (uint32_t)a_desc0, a_hi = (uint32_t)(a_desc0 >> 32);
      // the bias enters through the tensor core: ones[128 x 16] * biasB[256 x 16]^T initialises the accumulator (fp16 value + residual)
      const uint64_t ones_desc = umma_smem_desc(smem_u32(wsm + kE1OnesOff), 128 * 16, 128);
      const uint32_t ones_lo = (uint32_t)ones_desc, ones_hi = (uint32_t)(ones_desc >> 32);
      mbar_wait(wbar, 0, 42);
      uint32_t ws = 0, it = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++ws) {
        const int stage = ws % kE1Stages;
        mbar_wait(&full[stage], (ws / kE1Stages) & 1, 43);
        const uint32_t a_lo = a_lo0 + (uint32_t)(stage * (kE1WinBAl >> 4));
#pragma unroll 1
        for (int tt = 0; tt < kE1TilesPerUnit; ++tt, ++it) {
          const int acc = it & 1;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            mbar_wait(&tempty[2 * acc + hh], ((it >> 1) & 1) ^ 1, 44);
            tc_fence_after();
            const uint32_t d = tmem_base + acc * 256 + hh * 128;
            const uint32_t b_lo = b_lo0 + (uint32_t)(hh * 128);   // rows [128 hh, 128 hh + 128) of every B image
            umma_f16_lohi(d, ones_lo, ones_hi, b_lo + (uint32_t)(kE1BiasOff >> 4), b_hi, idesc, 0u);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)   // window column 0 is 16u - 1: tap kw starts kw columns in; tile tt starts at row 8 tt
              umma_f16_lohi(d, a_lo + (uint32_t)(kw * kE1Blocks + 8 * tt), a_hi, b_lo + (uint32_t)(kw * (8192 >> 4)), b_hi, idesc, 1u);
            umma_commit(&tfull[2 * acc + hh]);
          }
        }
        umma_commit(&empty[stage]);
      }
    }
  } else if (warp < kE1EpiWarps) {
    // ===================== epilogue =====================
    const int q = warp & 3;          // TMEM lane quarter
    const int h = (warp >> 2) & 1;   // channel half: channels 16h .. 16h+15
    const int grp = warp >> 3;       // accumulator / tile parity this warp serves
    const int g = 4 * q + (lane >> 3);   // column within the unit
    const int i = lane & 7;              // time block within the tile
    const int odd = g & 1;               // feature parity: even column keeps channels 16h..16h+7, odd column 16h+8..16h+15
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const long long gc = 16ll * u + g;
      const long long n = gc / kE1Cols;
      const int fpp = (int)(gc - n * kE1Cols);
      const bool valid = (fpp >= 2) && (fpp <= kF + 1) && (n < p.n_utts);
      // the lane pair (f'' even, f'' + 1) is one pooled feature column fo = (f'' - 2) / 2
      uint16_t* ocol = p.out + (long long)(2 * h + odd) * p.out_plane_elems + ((n * p.out_cols + ((fpp - 2) >> 1) + 1) * (long long)p.out_rs) * 8;
      for (int tt = 0; tt < kE1TilesPerUnit; ++tt, ++it) {
        if ((int)(it & 1) != grp) continue;
        const int acc = grp;
        const int tb = 8 * tt + i;
        mbar_wait(&tfull[2 * acc], (it >> 1) & 1, 45);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + acc * 256 + 16 * h;
        uint32_t pk[4][4];
        // TMEM reads are software-pipelined: the loads of step k+1 are in flight while step k is converted
        float a[2][16], b[2][16];
        tmem_ld_32x16(taddr, a[0]);
        tmem_ld_32x16(taddr + 32, b[0]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // pooled time step within the block: conv time offsets jj = 2k, 2k+1
          tmem_ld_wait();
          if (k & 1) {  // all TMEM reads of this warp from half k >> 1 are done: hand it back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[2 * acc + (k >> 1)]);
          }
          if (k == 1) {
            mbar_wait(&tfull[2 * acc + 1], (it >> 1) & 1, 46);
            tc_fence_after();
          }
          if (k < 3) {
            tmem_ld_32x16(taddr + (2 * k + 2) * 32, a[(k + 1) & 1]);
            tmem_ld_32x16(taddr + (2 * k + 3) * 32, b[(k + 1) & 1]);
          }
          const float* av = a[k & 1];
          const float* bv = b[k & 1];
          float o[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) o[c] = relu_nan(av[c]) + relu_nan(bv[c]);   // the bias is already in the accumulator
          float v[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float send = odd ? o[c] : o[c + 8];
            const float mine = odd ? o[c + 8] : o[c];
            v[c] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) pk[k][c] = pack_act2(v[2 * c], v[2 * c + 1]);
        }
        if (valid) {
          // pooled step 4tb + k: parity plane k & 1 (4 planes further), rows 2tb + 1 (k < 2) and 2tb + 2: the two rows of a
          // parity are 32 contiguous bytes, the 8 lanes of a column 256
#pragma unroll
          for (int par = 0; par < 2; ++par) {
            uint16_t* dst = ocol + (long long)(par * 4) * p.out_plane_elems + (2 * tb + 1) * 8;
            st_global_v4(dst, pk[par][0], pk[par][1], pk[par][2], pk[par][3]);
            st_global_v4(dst + 8, pk[par + 2][0], pk[par + 2][1], pk[par + 2][2], pk[par + 2][3]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kE1EpiWarps + 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_cae_enc1_tc(const float* x, int64_t sn, int64_t st, int64_t sf, int n_utts, const float* norm_mean, const float* norm_std, uint16_t* xt,
                       const uint16_t* wpack, const float* bias_quarter, ActBuf out, int out_cols, int num_sms, cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  if (sf == 1) {   // feature-contiguous storage: transpose through shared memory (xt_prep.cuh)
    if (norm_mean != nullptr)
      xt_prep_transpose_kernel<false, true><<<dim3((kF + 31) / 32, n_utts), 256, 0, stream>>>(x, sn, st, kE1Cols, 2, kE1Lead, norm_mean, norm_std, xt);
    else
      xt_prep_transpose_kernel<false, false><<<dim3((kF + 31) / 32, n_utts), 256, 0, stream>>>(x, sn, st, kE1Cols, 2, kE1Lead, nullptr, nullptr, xt);
  } else {
    const long long total = (long long)n_utts * kF * kE1Blocks;
    cae_enc1_prep_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, stream>>>(x, sn, st, sf, total, norm_mean, norm_std, xt);
  }
  DFS_LAUNCH_CHECK();
  static bool configured[32] = {false};
  if (dfs_first_use_on_device(configured))
    DFS_CUDA_CHECK(cudaFuncSetAttribute(cae_enc1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kE1SmemB));
  Enc1TcParams p{};
  p.xt = xt;
  p.wpack = wpack;
  (void)bias_quarter;   // rides in wpack's bias image
  p.n_units = (int)ceil_div64((long long)n_utts * kE1Cols, kE1UnitCols);
  p.n_utts = n_utts;
  p.out = out.ptr;
  p.out_plane_elems = out.plane_elems();
  p.out_rs = out.RS;
  p.out_cols = out_cols;
  const int grid = p.n_units < num_sms ? p.n_units : num_sms;
  cae_enc1_tc_kernel<<<grid, kE1Threads, kE1SmemB, stream>>>(p);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

}  // namespace dfs
