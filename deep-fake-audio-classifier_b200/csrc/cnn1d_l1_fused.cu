// cnn1d_l1_fused.cu -- first layer of the 1D-CNN reading the caller's fp32 features directly:
//   x.transpose(1,2) -> Conv1d(180,32,k=3,p=1) + BatchNorm1d + ReLU        /root/reference/src/model_cnn1d.py:16-19,38-39
//
// The template path (cnn1d_tc.cu) first rewrites the input as fp16 FT8 planes (cnn1d_prep_kernel) and then TMA-loads
// them; on a dense [N,321,180] input that costs one extra HBM round trip of the largest tensor of the whole model
// (measured: prep 507 us + layer 1 165 us of an 843 us pass over 4,736 utterances).  Here twelve producer warps do the
// conversion in flight: coalesced 16-byte loads of the fp32 rows (a row of one utterance = 720 contiguous bytes),
// clamp + fp16, 16-byte st.shared straight into the SWIZZLE_NONE K-major image the MMA reads
//     stage[plane c8][utterance column 16][row 10][8 halfs],   plane stride 2576 B (+16 B: conflict-free 16-byte stores)
// followed by fence.proxy.async + mbarrier arrive.  GEMM view as in conv_tc.cuh MODE_3X1: a tile = 16 utterances x 8 time
// steps (M = 128), N = 32 output channels, K = 3 taps x 192 (180 features zero padded), a tap = +16 B on the A descriptor.
// The weights stay in the template's packing ([tap][24][64][8], upper 32 rows zero) and only rows 0..31 are read.
// Output: layer-1 activations in FT8 (planes 0..3 of the 8-plane buffer layer 2 reads; planes 4..7 stay zero).
// Used when the features are dense with the feature axis fastest (stride_f = 1, stride_t = 180, 16-byte aligned);
// any other layout takes the prep + template path.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "layout.cuh"

namespace dfs {

constexpr int kL1Rows = 10;                              // 8 time steps + 1 halo row each side
constexpr int kL1PlaneB = kColTile * kL1Rows * 16 + 16;  // 2576
constexpr int kL1Planes = 24;                            // 192 / 8
constexpr int kL1StageB = kL1Planes * kL1PlaneB;         // 61,824
constexpr int kL1Stages = 2;
constexpr int kL1WgtB = 3 * kL1Planes * 64 * 16;         // 73,728
constexpr int kL1Acc = 4;
constexpr int kL1N = 32;
constexpr int kL1Tiles = 41;                             // 328 / 8 time tiles per unit
constexpr int kL1StageOff = (kL1WgtB + 127) & ~127;
constexpr int kL1BarOff = kL1StageOff + kL1Stages * kL1StageB;
constexpr int kL1SmemB = kL1BarOff + 256;
constexpr int kL1ProdWarps = 12;                         // enough loads in flight to cover the L2 / HBM latency of the fp32 rows
constexpr int kL1MmaWarp = 4 + kL1ProdWarps;
constexpr int kL1Threads = 32 * (kL1MmaWarp + 2);        // warps 0-3 epilogue, 4-15 producers, 16 MMA issuer, 17 TMEM allocator

struct L1FusedParams {
  const float* x;          // dense [n][321][180] fp32
  long long sn;            // utterance stride in elements
  const uint16_t* wpack;   // [tap 3][24][64][8] fp16
  float bias[32];
  int n_units;             // 16-utterance column tiles
  int n_utts;
  uint16_t* out;           // FT8, 8 planes, RS rows per column, column 1 + n
  long long out_plane_elems;
  int out_rs;
  int pf;                  // L2 prefetch distance in tiles
};

__global__ void __launch_bounds__(kL1Threads, 1) cnn1d_l1_fused_kernel(const __grid_constant__ L1FusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* wsm = smem;
  uint8_t* stage0 = smem + kL1StageOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kL1BarOff);
  uint64_t* full = bars;                   // [stages]  producers -> MMA (one arrival per producer warp)
  uint64_t* empty = bars + kL1Stages;      // [stages]  MMA -> producers
  uint64_t* tfull = empty + kL1Stages;     // [acc]     MMA -> epilogue
  uint64_t* tempty = tfull + kL1Acc;       // [acc]     epilogue -> MMA (4 warps)
  uint64_t* wbar = tempty + kL1Acc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // zero both stages once: plane 23, the upper half of plane 22 and the +16 B pads are never written again
  for (int i = threadIdx.x; i < kL1Stages * kL1StageB / 16; i += kL1Threads) reinterpret_cast<uint4*>(stage0)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == kL1MmaWarp && lane == 0) {
    for (int i = 0; i < kL1Stages; ++i) { mbar_init(&full[i], kL1ProdWarps); mbar_init(&empty[i], 1); }
    for (int i = 0; i < kL1Acc; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == kL1MmaWarp + 1) {
    tmem_alloc(tmem_slot, kL1Acc * kL1N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 4 && warp < kL1MmaWarp) {
    // ===================== producers: fp32 rows -> fp16 K-major stage =====================
    const int pt = threadIdx.x - 128;
    if (pt == 0) {
      mbar_arrive_expect_tx(wbar, kL1WgtB);
      for (int off = 0; off < kL1WgtB; off += 16384) {
        const int bytes = (kL1WgtB - off) < 16384 ? (kL1WgtB - off) : 16384;
        bulk_g2s(wsm + off, reinterpret_cast<const uint8_t*>(p.wpack) + off, bytes, wbar);
      }
    }
    // L2 prefetch, kPf tiles ahead: thread pt < 16 asks for the 8 new rows (5,760 contiguous bytes) of its utterance, so the
    // loads below find their lines in L2 (~300 cycles) instead of paying the HBM latency once per round
    const int kPf = p.pf;
    auto prefetch_rows = [&](int u, int tt) {
      if (pt >= kColTile || u >= p.n_units) return;
      const long long gn = (long long)kColTile * u + pt;
      if (gn >= p.n_utts) return;
      const int t0 = 8 * tt, t1 = (8 * tt + 8) < kT ? (8 * tt + 8) : kT;
      if (t0 >= t1) return;
      const float* src = p.x + gn * p.sn + (long long)t0 * kF;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)((t1 - t0) * kF * 4)) : "memory");
    };
    for (int k = 0; k < kPf; ++k) prefetch_rows(blockIdx.x, k);
    uint32_t ws = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      for (int tt = 0; tt < kL1Tiles; ++tt, ++ws) {
        if (tt + kPf < kL1Tiles) prefetch_rows(u, tt + kPf);
        else prefetch_rows(u + gridDim.x, tt + kPf - kL1Tiles);
        const int stage = ws % kL1Stages;
        mbar_wait(&empty[stage], ((ws / kL1Stages) & 1) ^ 1, 51);
        uint8_t* sb = stage0 + stage * kL1StageB;
        // thread pt < 368 owns one (utterance column, feature chunk) pair and walks its 10 rows: the addresses are plain
        // increments (no per-item div / mod), consecutive threads read consecutive 32-byte chunks of a 720-byte row
        if (pt < kColTile * 23) {
          const int col = pt / 23, c8 = pt - col * 23;
          const long long gn = (long long)kColTile * u + col;
          const bool uvalid = gn < p.n_utts;
          const float* base = p.x + (uvalid ? gn : 0) * p.sn + 8 * c8;
          uint8_t* dst = sb + c8 * kL1PlaneB + col * (kL1Rows * 16);
          constexpr int U = 5;
#pragma unroll
          for (int r0 = 0; r0 < kL1Rows; r0 += U) {
            float4 lo[U], hi[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
              const int t = 8 * tt - 1 + r0 + k;
              lo[k] = make_float4(0.f, 0.f, 0.f, 0.f);
              hi[k] = lo[k];
              if (uvalid && t >= 0 && t < kT) {
                const float4* src = reinterpret_cast<const float4*>(base + (long long)t * kF);
                lo[k] = __ldg(src);
                if (c8 < 22) hi[k] = __ldg(src + 1);   // chunk 22 = features 176..179 + zero padding
              }
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
              // pack_act2 saturates both ways (F2FP.SATFINITE): no separate clamp instructions in this issue-sensitive loop
              const uint4 v = make_uint4(pack_act2(lo[k].x, lo[k].y), pack_act2(lo[k].z, lo[k].w), pack_act2(hi[k].x, hi[k].y), pack_act2(hi[k].z, hi[k].w));
              *reinterpret_cast<uint4*>(dst + (r0 + k) * 16) = v;
            }
          }
        }
        fence_proxy_async_smem();   // every lane: generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();               // then ONE arrival per warp (hundreds of single-thread arrivals on one mbarrier serialise: ~4 cycles each)
        if (lane == 0) mbar_arrive(&full[stage]);
      }
    }
  } else if (warp == kL1MmaWarp) {
    // ===================== MMA issuer =====================
    if (elect_one_sync()) {   // not `lane == 0`: see conv_tc.cuh
      constexpr uint32_t idesc = umma_idesc_f16(128, kL1N);
      const uint64_t b_desc0 = umma_smem_desc(smem_u32(wsm), 64 * 16, 128);        // K chunk stride = 64 rows of 16 B
      const uint32_t b_lo0 = (uint32_t)b_desc0, b_hi = (uint32_t)(b_desc0 >> 32);
      const uint64_t a_desc0 = umma_smem_desc(smem_u32(stage0), kL1PlaneB, kL1Rows * 16);
      const uint32_t a_lo0 = (uint32_t)a_desc0, a_hi = (uint32_t)(a_desc0 >> 32);
      mbar_wait(wbar, 0, 52);
      uint32_t ws = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        for (int tt = 0; tt < kL1Tiles; ++tt, ++ws) {
          const int stage = ws % kL1Stages, acc = ws % kL1Acc;
          mbar_wait(&full[stage], (ws / kL1Stages) & 1, 53);
          mbar_wait(&tempty[acc], ((ws / kL1Acc) & 1) ^ 1, 54);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)(stage * (kL1StageB >> 4));
#pragma unroll
          for (int tap = 0; tap < 3; ++tap) {
#pragma unroll
            for (int kk = 0; kk < kL1Planes / 2; ++kk) {
              const uint32_t a_off = (uint32_t)((2 * kk * kL1PlaneB + tap * 16) >> 4);
              const uint32_t b_off = (uint32_t)(((tap * kL1Planes + 2 * kk) * 64 * 16) >> 4);
              umma_f16_lohi(tmem_base + acc * kL1N, a_lo + a_off, a_hi, b_lo0 + b_off, b_hi, idesc, (tap | kk) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&tfull[acc]);
          umma_commit(&empty[stage]);
        }
      }
    }
  } else if (warp < 4) {
    // ===================== epilogue: bias + ReLU -> FT8 planes 0..3 =====================
    const int q = warp;
    const int r = 32 * q + lane, g = r >> 3, i = r & 7;
    uint32_t ws = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const long long n = (long long)kColTile * u + g;
      const bool colvalid = n < p.n_utts;
      for (int tt = 0; tt < kL1Tiles; ++tt, ++ws) {
        const int acc = ws % kL1Acc;
        mbar_wait(&tfull[acc], (ws / kL1Acc) & 1, 55);
        tc_fence_after();
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + acc * kL1N, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) pk[c >> 1] = pack_act2(relu_nan(v[c] + p.bias[c]), relu_nan(v[c + 1] + p.bias[c + 1]));
        const int tp = 1 + 8 * tt + i;
        if (colvalid && tp <= kT) {
          uint16_t* dst = p.out + ((1 + n) * (long long)p.out_rs + tp) * 8;
#pragma unroll
          for (int k = 0; k < 4; ++k) st_global_v4(dst + k * p.out_plane_elems, pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kL1MmaWarp + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kL1Acc * kL1N);
  }
}

bool cnn1d_l1_fused_supported(const float* x, int64_t sn, int64_t st, int64_t sf) {
  return sf == 1 && st == kF && (sn % 4) == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0;
}

int launch_cnn1d_l1_fused(const float* x, int64_t sn, int n_utts, const uint16_t* wpack, const float* bias, ActBuf out, int num_sms, cudaStream_t stream) {
  if (n_utts <= 0) return DFS_OK;
  static bool configured[32] = {false};
  if (dfs_first_use_on_device(configured))
    DFS_CUDA_CHECK(cudaFuncSetAttribute(cnn1d_l1_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kL1SmemB));
  L1FusedParams p{};
  p.x = x;
  p.sn = sn;
  p.wpack = wpack;
  for (int i = 0; i < 32; ++i) p.bias[i] = bias[i];
  p.n_units = (n_utts + kColTile - 1) / kColTile;
  p.n_utts = n_utts;
  p.out = out.ptr;
  p.out_plane_elems = out.plane_elems();
  p.out_rs = out.RS;
  p.pf = 1;
  if (const char* e = getenv("DFS_L1_PF")) p.pf = atoi(e);
  const int grid = p.n_units < num_sms ? p.n_units : num_sms;
  cnn1d_l1_fused_kernel<<<grid, kL1Threads, kL1SmemB, stream>>>(p);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

}  // namespace dfs
