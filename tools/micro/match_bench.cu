// micro-benchmark: cost of finding the lanes with the same 8-bit digit -- 8 ballots vs match.any.sync (sm_100a)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/match_bench tools/micro/match_bench.cu
// B200 result: 8 ballots 362 G keys/s chip-wide, match.any 154 G keys/s
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t peers_ballot(uint32_t d) {
  uint32_t peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (d >> b) & 1u;
    const uint32_t bal = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? bal : ~bal;
  }
  return peers;
}

template <int MODE>
__global__ void bench(const uint32_t* in, uint32_t* out, int iters, long long* cycles) {
  uint32_t v = in[blockIdx.x * blockDim.x + threadIdx.x];
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t d = (v >> (i & 15)) & 0xffu;
    uint32_t p;
    if (MODE == 0) p = peers_ballot(d);
    else p = __match_any_sync(0xffffffffu, d);
    acc += __popc(p) + (p & 1u);
    v = v * 1664525u + 1013904223u;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  uint32_t *in, *out;
  long long* cyc;
  cudaMalloc(&in, blocks * threads * 4);
  cudaMalloc(&out, blocks * threads * 4);
  cudaMallocManaged(&cyc, 8);
  uint32_t* h = new uint32_t[blocks * threads];
  for (int i = 0; i < blocks * threads; ++i) h[i] = (uint32_t)rand() * 2654435761u;
  cudaMemcpy(in, h, blocks * threads * 4, cudaMemcpyHostToDevice);
  for (int mode = 0; mode < 2; ++mode) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) bench<0><<<blocks, threads>>>(in, out, iters, cyc);
      else bench<1><<<blocks, threads>>>(in, out, iters, cyc);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_ops = (double)blocks * threads / 32 * iters;
    printf("%s: %.3f ms, %.1f cycles/iter (one warp, incl. loop overhead), %.2f G warp-ops/s -> %.2f G keys/s chip-wide\n",
           mode == 0 ? "8 ballots " : "match.any ", ms, (double)*cyc / iters, warp_ops / ms / 1e6, warp_ops * 32 / ms / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
