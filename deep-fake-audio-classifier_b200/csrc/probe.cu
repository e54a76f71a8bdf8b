// probe.cu -- bring-up probes for the two hardware contracts the conv kernels rely on
// (tests/test_gpu_probes.py).  They exercise exactly the helpers of common.cuh:
//   probe_umma        one tcgen05.mma tile through SWIZZLE_NONE K-major descriptors, with the start
//                     address shifted by an arbitrary number of 16-byte rows and an arbitrary
//                     8-row-group stride (SBO) -- the addressing the 3x3 taps use.
//   probe_tma_window  one 3-D TMA box load of an FT8 activation window, dumped back to global.
//   probe_umma_bench  tcgen05.mma issue / operand-fetch cost for the addressing patterns of the conv kernels
//   probe_tmem_ld_bench  tcgen05.ld read-out rate per SM for the shapes and warp counts an epilogue can use
// Built into lib/libdfs_b200_probes.so (test / measurement only; csrc/probes.h), NOT into the product library.
#include "common.cuh"
#include "kernels.h"
#include "probes.h"
#include "tmem_ld_shapes.cuh"

#include <stdarg.h>

#include <vector>

// the probes library carries its own copies of the error plumbing common.cuh declares
static thread_local char g_probe_err[1024] = "";
void dfs_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_probe_err, sizeof(g_probe_err), fmt, ap);
  va_end(ap);
}
void dfs_count_launch(int) {}
extern "C" const char* dfs_probe_last_error(void) { return g_probe_err; }

namespace dfs {

// A: [rows_a][K] row-major bf16 bits, B: [N][K] row-major.  Staged as [K/8][rows][8] planes.
// D row r = 8g+i reads staged A row (row_shift + g*group_rows + i).
__global__ void __launch_bounds__(128) probe_umma_kernel(const uint16_t* __restrict__ a, const uint16_t* __restrict__ b, int rows_a, int N,
                                                          int K, int row_shift, int group_rows, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint16_t* sa = reinterpret_cast<uint16_t*>(smem);
  uint16_t* sb = sa + (size_t)(K / 8) * rows_a * 8;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (((size_t)(K / 8) * (rows_a + N) * 16 + 15) & ~(size_t)15));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int idx = threadIdx.x; idx < rows_a * K; idx += blockDim.x) {
    const int r = idx / K, k = idx - r * K;
    sa[((size_t)(k >> 3) * rows_a + r) * 8 + (k & 7)] = a[idx];
  }
  for (int idx = threadIdx.x; idx < N * K; idx += blockDim.x) {
    const int r = idx / K, k = idx - r * K;
    sb[((size_t)(k >> 3) * N + r) * 8 + (k & 7)] = b[idx];
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    for (int kk = 0; kk < K / 16; ++kk) {
      const uint64_t da = umma_smem_desc(smem_u32(sa) + (2 * kk) * rows_a * 16 + row_shift * 16, rows_a * 16, group_rows * 16);
      const uint64_t db = umma_smem_desc(smem_u32(sb) + (2 * kk) * N * 16, N * 16, 128);
      umma_bf16(tmem_base, da, db, idesc, kk != 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0, 9);
  tc_fence_after();
  const int r = 32 * warp + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(32 * warp) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (c0 + c < N) out[(size_t)r * N + c0 + c] = v[c];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int probe_umma(const uint16_t* a, const uint16_t* b, int rows_a, int n, int k, int row_shift, int group_rows, float* out,
               cudaStream_t stream) {
  DFS_REQUIRE(a && b && out, DFS_ERR_INVALID, "probe_umma: NULL argument");
  DFS_REQUIRE(n % 32 == 0 && n >= 32 && n <= 256 && k % 16 == 0 && k >= 16, DFS_ERR_INVALID, "probe_umma: bad n/k");
  DFS_REQUIRE(group_rows >= 8 && row_shift >= 0 && row_shift + 15 * group_rows + 8 <= rows_a, DFS_ERR_INVALID,
              "probe_umma: window does not fit rows_a");
  const size_t bytes = (((size_t)(k / 8) * (rows_a + n) * 16 + 15) & ~(size_t)15) + 64;
  DFS_REQUIRE(bytes <= 200 * 1024, DFS_ERR_INVALID, "probe_umma: operands too large for shared memory");
  DFS_CUDA_CHECK(cudaFuncSetAttribute(probe_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  probe_umma_kernel<<<1, 128, bytes, stream>>>(a, b, rows_a, n, k, row_shift, group_rows, out);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

// Issue-rate / operand-fetch micro-benchmark: one thread issues `iters` rounds of `nmma` MMAs
// (M=128, N=n, K=16) whose A start addresses walk a caller-given list of byte offsets (e.g. the 9
// tap offsets of a conv tile), commits once per round and waits; reports SM cycles per MMA.
// Operand CONTENT is irrelevant (shared memory is zero-filled), only the addressing is timed.
constexpr int kBenchMaxMma = 96;
struct UmmaBenchParams {
  int n, nmma, iters;
  int n_acc;               // accumulators used in rotation (MMA i of a round targets accumulator i % n_acc); 1 = one dependent chain
  uint32_t a_off[kBenchMaxMma];      // byte offsets of the A start address per MMA of a round
  uint32_t b_off[kBenchMaxMma];      // byte offsets of the B start address per MMA
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t layout;         // descriptor layout_type field (0 none, 2 SW128, 4 SW64, 6 SW32)
  uint32_t use_base_offset;
};

__global__ void __launch_bounds__(128) probe_umma_bench_kernel(const __grid_constant__ UmmaBenchParams p, long long* __restrict__ cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int A_BYTES = 96 * 1024, B_BYTES = 64 * 1024;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if ((threadIdx.x >> 5) == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, p.n);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + A_BYTES);
    uint64_t* ad = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES + 64);   // descriptor tables in shared memory
    uint64_t* bd = ad + kBenchMaxMma;
    // accumulator of MMA i = i mod n_acc (n_acc is 1, 2 or 4): a mask and a shift, so that the issue loop stays as short as the conv
    // kernels' (an integer division per MMA made the ISSUING THREAD the bottleneck: 148 cycles per MMA for every N)
    const uint32_t accmask = (uint32_t)p.n_acc - 1u;
    const uint32_t nshift = (uint32_t)__ffs(p.n) - 1u;
    const int nacc = p.n_acc;
    for (int i = 0; i < p.nmma; ++i) {
      const uint32_t aa = a0 + p.a_off[i], bb = b0 + p.b_off[i];
      ad[i] = umma_smem_desc(aa, p.a_lbo, p.a_sbo) | ((uint64_t)p.layout << 61) | (p.use_base_offset ? ((uint64_t)((aa >> 7) & 7) << 49) : 0ull);
      bd[i] = umma_smem_desc(bb, p.b_lbo, p.b_sbo) | ((uint64_t)p.layout << 61) | (p.use_base_offset ? ((uint64_t)((bb >> 7) & 7) << 49) : 0ull);
    }
    uint32_t phase = 0;
    // warm-up round
    for (int i = 0; i < p.nmma; ++i) umma_bf16(tmem_base + (((uint32_t)i & accmask) << nshift), ad[i], bd[i], idesc, i >= nacc);
    umma_commit(bar);
    mbar_wait(bar, phase, 11);
    phase ^= 1;
    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
#pragma unroll 4
      for (int i = 0; i < p.nmma; ++i) umma_bf16(tmem_base + (((uint32_t)i & accmask) << nshift), ad[i], bd[i], idesc, i >= nacc);
      umma_commit(bar);
      mbar_wait(bar, phase, 12);
      phase ^= 1;
    }
    const long long t1 = clock64();
    cycles_out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Lean issue loop: ONE fixed A / B descriptor pair held in registers, the MMAs of a round unrolled by four with compile-time
// accumulator offsets -- no shared-memory table loads, no index arithmetic between two tcgen05.mma.  What remains is the tensor
// pipe's own cost per MMA (the table-driven loop above spends ~10 dependent instructions of the issuing thread per MMA, which is
// of the same order as an N <= 128 MMA itself).
template <int NACC>
__global__ void __launch_bounds__(128) probe_umma_lean_kernel(const __grid_constant__ UmmaBenchParams p, long long* __restrict__ cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int A_BYTES = 96 * 1024, B_BYTES = 64 * 1024;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if ((threadIdx.x >> 5) == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x < 32 && elect_one_sync()) {   // elect, not `threadIdx.x == 0`: no per-MMA ELECT / BRA.U.ANY loop in the SASS
    const uint32_t idesc = umma_idesc_bf16(128, p.n);
    const uint64_t ad = umma_smem_desc(smem_u32(smem) + p.a_off[0], p.a_lbo, p.a_sbo) | ((uint64_t)p.layout << 61);
    const uint64_t bd = umma_smem_desc(smem_u32(smem + A_BYTES) + p.b_off[0], p.b_lbo, p.b_sbo) | ((uint64_t)p.layout << 61);
    const uint32_t n = (uint32_t)p.n;
    const int groups = p.nmma / 4;
    uint32_t phase = 0;
    long long t0 = 0;
    for (int it = -1; it < p.iters; ++it) {     // round -1 = warm-up
      if (it == 0) t0 = clock64();
#pragma unroll
      for (int u = 0; u < 4; ++u) umma_bf16(tmem_base + (uint32_t)(u % NACC) * n, ad, bd, idesc, u >= NACC ? 1u : 0u);
      for (int g = 1; g < groups; ++g) {
#pragma unroll
        for (int u = 0; u < 4; ++u) umma_bf16(tmem_base + (uint32_t)(u % NACC) * n, ad, bd, idesc, 1u);
      }
      umma_commit(bar);
      mbar_wait(bar, phase, 15);
      phase ^= 1;
    }
    cycles_out[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// The same loop on a CTA pair (cluster of 2): the leader issues tcgen05.mma.cta_group::2 (M = 256, each CTA supplies its 128 rows of A
// and n/2 rows of B from the same shared-memory offsets) and commits to the barrier at the same offset in both CTAs.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe_umma_bench_pair_kernel(const __grid_constant__ UmmaBenchParams p,
                                                                                            long long* __restrict__ cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int A_BYTES = 96 * 1024, B_BYTES = 64 * 1024;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if ((threadIdx.x >> 5) == 0) {
    tmem_alloc_pair(tmem_slot, 256);
    tmem_relinquish_pair();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(256, p.n);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + A_BYTES);
    uint64_t* ad = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES + 64);
    uint64_t* bd = ad + kBenchMaxMma;
    for (int i = 0; i < p.nmma; ++i) {
      const uint32_t aa = a0 + p.a_off[i], bb = b0 + p.b_off[i];
      ad[i] = umma_smem_desc(aa, p.a_lbo, p.a_sbo) | ((uint64_t)p.layout << 61);
      bd[i] = umma_smem_desc(bb, p.b_lbo, p.b_sbo) | ((uint64_t)p.layout << 61);
    }
    uint32_t phase = 0;
    for (int i = 0; i < p.nmma; ++i)
      umma_f16_lohi_pair(tmem_base, (uint32_t)ad[i], (uint32_t)(ad[i] >> 32), (uint32_t)bd[i], (uint32_t)(bd[i] >> 32), idesc, i != 0);
    umma_commit_pair(bar);
    mbar_wait(bar, phase, 13);
    phase ^= 1;
    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
#pragma unroll 4
      for (int i = 0; i < p.nmma; ++i)
        umma_f16_lohi_pair(tmem_base, (uint32_t)ad[i], (uint32_t)(ad[i] >> 32), (uint32_t)bd[i], (uint32_t)(bd[i] >> 32), idesc, i != 0);
      umma_commit_pair(bar);
      mbar_wait(bar, phase, 14);
      phase ^= 1;
    }
    const long long t1 = clock64();
    cycles_out[0] = t1 - t0;
  }
  tc_fence_before();
  cluster_sync_all();
  if ((threadIdx.x >> 5) == 0) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 256);
  }
}

int probe_umma_bench(int n, int nmma, int iters, int n_acc, const uint32_t* a_off, const uint32_t* b_off, uint32_t a_lbo, uint32_t a_sbo,
                     uint32_t b_lbo, uint32_t b_sbo, uint32_t layout, uint32_t use_base_offset, long long* cycles_host, cudaStream_t stream) {
  DFS_REQUIRE(n % 16 == 0 && n >= 16 && n <= 256 && nmma >= 1 && nmma <= kBenchMaxMma && iters >= 1, DFS_ERR_INVALID, "probe_umma_bench: bad argument");
  DFS_REQUIRE((n_acc == 1 || n_acc == 2 || n_acc == 4) && n_acc * n <= 512 && (n & (n - 1)) == 0 && (!(use_base_offset & 2) || n_acc == 1),
              DFS_ERR_INVALID, "probe_umma_bench: n a power of two, n_acc in {1, 2, 4} accumulators of n columns within the 512 TMEM columns (pairs: n_acc = 1)");
  UmmaBenchParams p{};
  p.n = n; p.nmma = nmma; p.iters = iters; p.n_acc = n_acc;
  for (int i = 0; i < nmma; ++i) { p.a_off[i] = a_off[i]; p.b_off[i] = b_off[i]; }
  p.a_lbo = a_lbo; p.a_sbo = a_sbo; p.b_lbo = b_lbo; p.b_sbo = b_sbo; p.layout = layout; p.use_base_offset = use_base_offset & 1;
  long long* d = nullptr;
  DFS_CUDA_CHECK(cudaMalloc(&d, 8));
  const int smem = 160 * 1024 + 64 + 2 * kBenchMaxMma * 8;
  DFS_CUDA_CHECK(cudaFuncSetAttribute(probe_umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (use_base_offset & 4) {   // bit 2: lean issue loop (fixed descriptors in registers, nmma a multiple of 4)
    DFS_REQUIRE(nmma % 4 == 0, DFS_ERR_INVALID, "probe_umma_bench: the lean loop issues the MMAs in groups of four");
    if (n_acc == 1) {
      DFS_CUDA_CHECK(cudaFuncSetAttribute(probe_umma_lean_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      probe_umma_lean_kernel<1><<<1, 128, smem, stream>>>(p, d);
    } else if (n_acc == 2) {
      DFS_CUDA_CHECK(cudaFuncSetAttribute(probe_umma_lean_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      probe_umma_lean_kernel<2><<<1, 128, smem, stream>>>(p, d);
    } else {
      DFS_CUDA_CHECK(cudaFuncSetAttribute(probe_umma_lean_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      probe_umma_lean_kernel<4><<<1, 128, smem, stream>>>(p, d);
    }
  } else if (use_base_offset & 2) {   // bit 1: run on a CTA pair (cta_group::2); n is the N of the pair's MMA
    DFS_CUDA_CHECK(cudaFuncSetAttribute(probe_umma_bench_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_umma_bench_pair_kernel<<<2, 128, smem, stream>>>(p, d);
  } else {
    probe_umma_bench_kernel<<<1, 128, smem, stream>>>(p, d);
  }
  dfs_count_launch();
  cudaError_t e = cudaStreamSynchronize(stream);
  if (e == cudaSuccess) e = cudaMemcpy(cycles_host, d, 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  DFS_REQUIRE(e == cudaSuccess, DFS_ERR_CUDA, "probe_umma_bench: %s", cudaGetErrorString(e));
  return DFS_OK;
}

__global__ void __launch_bounds__(128) probe_tma_kernel(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int bytes,
                                                         uint16_t* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ((bytes + 127) & ~127));
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, bytes);
    tma_load_3d(smem, &tmap, c0, c1, 0, bar);
  }
  mbar_wait(bar, 0, 10);
  for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) out[i] = reinterpret_cast<const uint16_t*>(smem)[i];
}

int probe_tma_window(const uint16_t* act, int planes, int RS, int64_t ncols, int wrows, int row0, int col0, uint16_t* out,
                     cudaStream_t stream) {
  DFS_REQUIRE(act && out, DFS_ERR_INVALID, "probe_tma_window: NULL argument");
  ActBuf a{const_cast<uint16_t*>(act), planes, RS, ncols};
  CUtensorMap tmap;
  DFS_PROPAGATE(make_act_tensor_map(&tmap, a, wrows, kColTile + 2, planes));
  const int bytes = planes * (kColTile + 2) * wrows * 16;
  const size_t smem = ((bytes + 127) & ~127) + 64;
  DFS_CUDA_CHECK(cudaFuncSetAttribute(probe_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_tma_kernel<<<1, 128, smem, stream>>>(tmap, row0 * 8, col0, bytes, out);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

// ---- TMEM read-out rate --------------------------------------------------------------------------------------------
// `nwarps` warps (4, 8 or 16) of ONE CTA read their lane quadrant (warp % 4) of the 512 allocated columns with one
// tcgen05.ld shape, `lds_per_wait` loads in flight before each tcgen05.wait::ld, `iters` rounds.  Reported: SM cycles between
// two block-wide barriers, and the bytes moved (lanes x columns x 4 B per load); content is whatever TMEM holds.
//   shape ids: 0..4 = 32x32b .x8 .x16 .x32 .x64 .x128 | 5..7 = 16x256b .x4 .x8 .x16 | 8..10 = 16x128b .x8 .x16 .x32
template <int SHAPE>
__device__ __forceinline__ uint32_t tmem_ld_shape(uint32_t taddr) {
  constexpr int NR = SHAPE == 0 ? 8 : SHAPE == 1 ? 16 : SHAPE == 2 ? 32 : SHAPE == 3 ? 64 : SHAPE == 4 ? 128 : SHAPE == 5 ? 16 : SHAPE == 6 ? 32
                     : SHAPE == 7 ? 64 : SHAPE == 8 ? 16 : SHAPE == 9 ? 32 : 64;
  uint32_t r[NR];
  if constexpr (SHAPE == 0) tmem_ld_32x32b_x8(taddr, r);
  else if constexpr (SHAPE == 1) tmem_ld_32x32b_x16(taddr, r);
  else if constexpr (SHAPE == 2) tmem_ld_32x32b_x32(taddr, r);
  else if constexpr (SHAPE == 3) tmem_ld_32x32b_x64(taddr, r);
  else if constexpr (SHAPE == 4) tmem_ld_32x32b_x128(taddr, r);
  else if constexpr (SHAPE == 5) tmem_ld_16x256b_x4(taddr, r);
  else if constexpr (SHAPE == 6) tmem_ld_16x256b_x8(taddr, r);
  else if constexpr (SHAPE == 7) tmem_ld_16x256b_x16(taddr, r);
  else if constexpr (SHAPE == 8) tmem_ld_16x128b_x8(taddr, r);
  else if constexpr (SHAPE == 9) tmem_ld_16x128b_x16(taddr, r);
  else tmem_ld_16x128b_x32(taddr, r);
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < NR; ++i) x ^= r[i];
  return x;
}
// columns one load of the shape spans, lanes it reads
__host__ __device__ constexpr int tmem_shape_cols(int s) {
  return s == 0 ? 8 : s == 1 ? 16 : s == 2 ? 32 : s == 3 ? 64 : s == 4 ? 128 : s == 5 ? 32 : s == 6 ? 64 : s == 7 ? 128 : s == 8 ? 32 : s == 9 ? 64 : 128;
}
__host__ __device__ constexpr int tmem_shape_lanes(int s) { return s <= 4 ? 32 : 16; }

template <int SHAPE>
__global__ void __launch_bounds__(SHAPE == 4 ? 256 : 512) probe_tmem_ld_kernel(int iters, int lds_per_wait, long long* __restrict__ cycles_out,
                                                             uint32_t* __restrict__ sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_slot + ((uint32_t)(32 * (warp & 3)) << 16);
  constexpr int COLS = tmem_shape_cols(SHAPE);
  uint32_t acc = 0, col = (uint32_t)((warp >> 2) * COLS) & 511u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int j = 0; j < lds_per_wait; ++j) {
      acc ^= tmem_ld_shape<SHAPE>(base + col);
      col = (col + COLS) & 511u;                 // COLS divides 512: a load never runs past the allocation
    }
    tmem_ld_wait();
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
  if (acc == 0x9e3779b9u) sink[0] = acc;         // keeps the loaded registers alive
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_slot, 512);
  }
}

template <int SHAPE>
static void launch_tmem_ld(int blocks, int nwarps, int iters, int lpw, long long* cyc, uint32_t* sink, cudaStream_t stream) {
  probe_tmem_ld_kernel<SHAPE><<<blocks, nwarps * 32, 0, stream>>>(iters, lpw, cyc, sink);
}

int probe_tmem_ld_bench(int shape, int nwarps, int blocks, int iters, int lds_per_wait, long long* cycles_host, long long* bytes_per_block_host,
                        cudaStream_t stream) {
  DFS_REQUIRE(shape >= 0 && shape <= 10 && (nwarps == 4 || nwarps == 8 || nwarps == 16) && blocks >= 1 && blocks <= 148 && iters >= 1 &&
                  lds_per_wait >= 1 && cycles_host && bytes_per_block_host,
              DFS_ERR_INVALID, "probe_tmem_ld_bench: bad argument");
  DFS_REQUIRE(shape != 4 || nwarps <= 8, DFS_ERR_INVALID, "probe_tmem_ld_bench: 32x32b.x128 needs 128+ registers per thread: at most 8 warps");
  long long* d = nullptr;
  uint32_t* sink = nullptr;
  DFS_CUDA_CHECK(cudaMalloc(&d, 8 * (size_t)blocks));
  DFS_CUDA_CHECK(cudaMalloc(&sink, 4));
  switch (shape) {
    case 0: launch_tmem_ld<0>(blocks, nwarps, iters, lds_per_wait, d, sink, stream); break;
    case 1: launch_tmem_ld<1>(blocks, nwarps, iters, lds_per_wait, d, sink, stream); break;
    case 2: launch_tmem_ld<2>(blocks, nwarps, iters, lds_per_wait, d, sink, stream); break;
    case 3: launch_tmem_ld<3>(blocks, nwarps, iters, lds_per_wait, d, sink, stream); break;
    case 4: launch_tmem_ld<4>(blocks, nwarps, iters, lds_per_wait, d, sink, stream); break;
    case 5: launch_tmem_ld<5>(blocks, nwarps, iters, lds_per_wait, d, sink, stream); break;
    case 6: launch_tmem_ld<6>(blocks, nwarps, iters, lds_per_wait, d, sink, stream); break;
    case 7: launch_tmem_ld<7>(blocks, nwarps, iters, lds_per_wait, d, sink, stream); break;
    case 8: launch_tmem_ld<8>(blocks, nwarps, iters, lds_per_wait, d, sink, stream); break;
    case 9: launch_tmem_ld<9>(blocks, nwarps, iters, lds_per_wait, d, sink, stream); break;
    default: launch_tmem_ld<10>(blocks, nwarps, iters, lds_per_wait, d, sink, stream); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  long long worst = 0;
  if (e == cudaSuccess) {
    std::vector<long long> h(blocks);
    e = cudaMemcpy(h.data(), d, 8 * (size_t)blocks, cudaMemcpyDeviceToHost);
    for (long long c : h) worst = c > worst ? c : worst;
  }
  cudaFree(d);
  cudaFree(sink);
  DFS_REQUIRE(e == cudaSuccess, DFS_ERR_CUDA, "probe_tmem_ld_bench: %s", cudaGetErrorString(e));
  *cycles_host = worst;
  *bytes_per_block_host = (long long)nwarps * iters * lds_per_wait * tmem_shape_lanes(shape) * tmem_shape_cols(shape) * 4;
  return DFS_OK;
}

}  // namespace dfs

// ---- C exports of the probes library (csrc/probes.h) ------------------------------------------------------------------
using namespace dfs;
extern "C" int dfs_probe_umma(const uint16_t* a_dev, const uint16_t* b_dev, int rows_a, int n, int k, int row_shift, int group_rows,
                              float* out_dev, void* stream) {
  return probe_umma(a_dev, b_dev, rows_a, n, k, row_shift, group_rows, out_dev, static_cast<cudaStream_t>(stream));
}
extern "C" int dfs_probe_tma_window(const uint16_t* act_dev, int planes, int rs, int64_t ncols, int wrows, int row0, int col0,
                                    uint16_t* out_dev, void* stream) {
  return probe_tma_window(act_dev, planes, rs, ncols, wrows, row0, col0, out_dev, static_cast<cudaStream_t>(stream));
}
extern "C" int dfs_probe_umma_bench(int n, int nmma, int iters, int n_acc, const uint32_t* a_off_host, const uint32_t* b_off_host, uint32_t a_lbo,
                                    uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, uint32_t layout, uint32_t use_base_offset,
                                    int64_t* cycles_host, void* stream) {
  long long c = 0;
  int st = probe_umma_bench(n, nmma, iters, n_acc, a_off_host, b_off_host, a_lbo, a_sbo, b_lbo, b_sbo, layout, use_base_offset, &c,
                            static_cast<cudaStream_t>(stream));
  if (cycles_host) *cycles_host = c;
  return st;
}
extern "C" int dfs_probe_tmem_ld_bench(int shape, int nwarps, int blocks, int iters, int lds_per_wait, int64_t* cycles_host,
                                       int64_t* bytes_per_block_host, void* stream) {
  long long c = 0, b = 0;
  int st = probe_tmem_ld_bench(shape, nwarps, blocks, iters, lds_per_wait, &c, &b, static_cast<cudaStream_t>(stream));
  if (cycles_host) *cycles_host = c;
  if (bytes_per_block_host) *bytes_per_block_host = b;
  return st;
}
