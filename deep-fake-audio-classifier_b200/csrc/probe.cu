// probe.cu -- bring-up probes for the two hardware contracts the conv kernels rely on
// (tests/test_gpu_probes.py).  They exercise exactly the helpers of common.cuh:
//   probe_umma        one tcgen05.mma tile through SWIZZLE_NONE K-major descriptors, with the start
//                     address shifted by an arbitrary number of 16-byte rows and an arbitrary
//                     8-row-group stride (SBO) -- the addressing the 3x3 taps use.
//   probe_tma_window  one 3-D TMA box load of an FT8 activation window, dumped back to global.
#include "common.cuh"
#include "kernels.h"

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
struct UmmaBenchParams {
  int n, nmma, iters;
  uint32_t a_off[40];      // byte offsets of the A start address per MMA of a round
  uint32_t b_off[40];      // byte offsets of the B start address per MMA
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
    tmem_alloc(tmem_slot, 256);
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
    uint64_t* bd = ad + 40;
    for (int i = 0; i < p.nmma; ++i) {
      const uint32_t aa = a0 + p.a_off[i], bb = b0 + p.b_off[i];
      ad[i] = umma_smem_desc(aa, p.a_lbo, p.a_sbo) | ((uint64_t)p.layout << 61) | (p.use_base_offset ? ((uint64_t)((aa >> 7) & 7) << 49) : 0ull);
      bd[i] = umma_smem_desc(bb, p.b_lbo, p.b_sbo) | ((uint64_t)p.layout << 61) | (p.use_base_offset ? ((uint64_t)((bb >> 7) & 7) << 49) : 0ull);
    }
    uint32_t phase = 0;
    // warm-up round
    for (int i = 0; i < p.nmma; ++i) umma_bf16(tmem_base, ad[i], bd[i], idesc, i != 0);
    umma_commit(bar);
    mbar_wait(bar, phase, 11);
    phase ^= 1;
    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
#pragma unroll 4
      for (int i = 0; i < p.nmma; ++i) umma_bf16(tmem_base, ad[i], bd[i], idesc, i != 0);
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
    tmem_dealloc(tmem_base, 256);
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
    uint64_t* bd = ad + 40;
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

int probe_umma_bench(int n, int nmma, int iters, const uint32_t* a_off, const uint32_t* b_off, uint32_t a_lbo, uint32_t a_sbo,
                     uint32_t b_lbo, uint32_t b_sbo, uint32_t layout, uint32_t use_base_offset, long long* cycles_host, cudaStream_t stream) {
  DFS_REQUIRE(n % 16 == 0 && n >= 16 && n <= 256 && nmma >= 1 && nmma <= 40 && iters >= 1, DFS_ERR_INVALID, "probe_umma_bench: bad argument");
  UmmaBenchParams p{};
  p.n = n; p.nmma = nmma; p.iters = iters;
  for (int i = 0; i < nmma; ++i) { p.a_off[i] = a_off[i]; p.b_off[i] = b_off[i]; }
  p.a_lbo = a_lbo; p.a_sbo = a_sbo; p.b_lbo = b_lbo; p.b_sbo = b_sbo; p.layout = layout; p.use_base_offset = use_base_offset & 1;
  long long* d = nullptr;
  DFS_CUDA_CHECK(cudaMalloc(&d, 8));
  const int smem = 160 * 1024 + 64 + 80 * 8;
  DFS_CUDA_CHECK(cudaFuncSetAttribute(probe_umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (use_base_offset & 2) {   // bit 1: run on a CTA pair (cta_group::2); n is the N of the pair's MMA
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

}  // namespace dfs
