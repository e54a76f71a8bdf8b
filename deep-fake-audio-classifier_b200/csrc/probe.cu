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
  DFS_PROPAGATE(make_act_tensor_map(&tmap, a, wrows));
  const int bytes = planes * (kColTile + 2) * wrows * 16;
  const size_t smem = ((bytes + 127) & ~127) + 64;
  DFS_CUDA_CHECK(cudaFuncSetAttribute(probe_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_tma_kernel<<<1, 128, smem, stream>>>(tmap, row0 * 8, col0, bytes, out);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

}  // namespace dfs
