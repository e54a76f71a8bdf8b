// synth.cu -- counter-based synthetic LFCC-like maps, generated on the device so that the
// benchmark pool never crosses PCIe and every rank / GPU count sees the same global data set
// (BASELINE.json north_star: "pinned synthetic feature tensors, replacing src/dataloaders.py's
// pickle path on the benchmark").
#include "common.cuh"
#include "kernels.h"

namespace dfs {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// element pair (2k, 2k+1) of utterance u <- Box-Muller on one 64-bit hash of (seed, u, k)
__global__ void __launch_bounds__(256) fill_features_kernel(float* __restrict__ out, long long n, long long first_utt, uint64_t seed, float sd) {
  constexpr long long PAIRS = (long long)kT * kF / 2;
  const long long total = n * PAIRS;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long u = i / PAIRS, k = i - u * PAIRS;
    const uint64_t h = splitmix64(splitmix64(seed ^ ((uint64_t)(first_utt + u) * 0xD1B54A32D192ED03ull)) + (uint64_t)k);
    const float u1 = ((float)(uint32_t)(h >> 40) + 1.0f) * (1.0f / 16777216.0f);  // (0, 1]
    const float u2 = (float)(uint32_t)((h >> 8) & 0xffffffu) * (1.0f / 16777216.0f);
    const float r = sd * sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    reinterpret_cast<float2*>(out)[i] = make_float2(r * c, r * s);
  }
}

int fill_features_device(float* out, int64_t n, int64_t first_utt, uint64_t seed, float sd, cudaStream_t stream) {
  DFS_REQUIRE(out != nullptr && n >= 0, DFS_ERR_INVALID, "dfs_fill_features: bad argument");
  if (n == 0) return DFS_OK;
  fill_features_kernel<<<148 * 8, 256, 0, stream>>>(out, n, first_utt, seed, sd);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

}  // namespace dfs
