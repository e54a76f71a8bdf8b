// eer.cu -- the metric half of the hot path on the device:
//   calculate_eer            /root/reference/scripts/evaluation.py:7-39 (== src/evaluation.py:12-48)
//   confusion_at_threshold   /root/reference/scripts/evaluation.py:42-56
//   normalise_01 + blend     /root/reference/src/predict_hybrid.py:81-85,149-151, src/ensemble.py:121
//
// EER = stable LSD radix sort (one-sweep: per-pass decoupled look-back, 8-bit digits, key =
// order-preserving integer image of the fp32/fp64 score, payload = original index | label<<31)
// followed by a prefix-count FAR/FRR sweep in IEEE fp64 and a (value, lowest index) arg-min.
// All kernels are HBM-bound streaming kernels: 16-byte vector accesses where the layout allows,
// one tile of 4096 keys per CTA, tiles dealt by an atomic ticket so look-back cannot deadlock.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace dfs {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096
constexpr uint32_t kFlagAgg = 1u << 30, kFlagIncl = 2u << 30, kValMask = (1u << 30) - 1;

// ---- order-preserving key transforms (-0.0 is canonicalised to +0.0: numpy treats them as ties) ----
__device__ __forceinline__ uint32_t to_key(float s) {
  uint32_t b = __float_as_uint(s);
  if (b == 0x80000000u) b = 0;
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ uint64_t to_key(double s) {
  uint64_t b = (uint64_t)__double_as_longlong(s);
  if (b == 0x8000000000000000ull) b = 0;
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ float from_key(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ double from_key(uint64_t k) {
  return __longlong_as_double((long long)((k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k));
}
template <typename K> struct ScoreOf;
template <> struct ScoreOf<uint32_t> { using type = float; };
template <> struct ScoreOf<uint64_t> { using type = double; };

// ---- pass 0: keys, payloads, all digit histograms, label count -------------------------------
template <typename K>
__global__ void __launch_bounds__(256) sort_prep_kernel(const typename ScoreOf<K>::type* __restrict__ scores,
                                                         const uint8_t* __restrict__ labels, long long n, K* __restrict__ keys,
                                                         uint32_t* __restrict__ pay, uint32_t* __restrict__ ghist /*[passes][256]*/,
                                                         unsigned long long* __restrict__ n_ones) {
  constexpr int PASSES = sizeof(K);
  __shared__ uint32_t hist[PASSES][256];
  for (int i = threadIdx.x; i < PASSES * 256; i += blockDim.x) (&hist[0][0])[i] = 0;
  __syncthreads();
  uint32_t ones = 0;
  // warp-uniform trip count so the match/ballot below always sees the full warp
  for (long long i0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += (long long)gridDim.x * blockDim.x) {
    const long long i = i0 + (threadIdx.x & 31);
    const bool ok = i < n;
    K k = 0;
    if (ok) {
      k = to_key(scores[i]);
      const uint32_t lab = labels[i] != 0;
      keys[i] = k;
      pay[i] = (uint32_t)i | (lab << 31);
      ones += lab;
    }
    // warp-aggregated: real score vectors share their high digits, plain atomics would serialise
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const uint32_t d = ok ? ((uint32_t)(k >> (8 * ps)) & 0xffu) : 0xffffffffu;
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      if (ok && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[ps][d], (uint32_t)__popc(peers));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PASSES * 256; i += blockDim.x) {
    const uint32_t v = (&hist[0][0])[i];
    if (v) atomicAdd(&ghist[i], v);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) ones += __shfl_xor_sync(0xffffffffu, ones, o);
  if ((threadIdx.x & 31) == 0 && ones) atomicAdd(n_ones, (unsigned long long)ones);
}

// exclusive scan of each pass's 256-bin histogram; skip[pass] = 1 when one bin holds every key
__global__ void __launch_bounds__(256) sort_scan_hist_kernel(uint32_t* __restrict__ ghist, uint32_t* __restrict__ skip, long long n) {
  __shared__ uint32_t s[256];
  const uint32_t v = ghist[blockIdx.x * 256 + threadIdx.x];
  s[threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    bool one_bin = false;
    for (int i = 0; i < 256; ++i) {
      const uint32_t c = s[i];
      if ((long long)c == n) one_bin = true;
      s[i] = run;
      run += c;
    }
    skip[blockIdx.x] = one_bin ? 1u : 0u;
  }
  __syncthreads();
  ghist[blockIdx.x * 256 + threadIdx.x] = s[threadIdx.x];
}

// ---- one radix pass ----------------------------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(kSortThreads) onesweep_pass_kernel(const K* __restrict__ kin, const uint32_t* __restrict__ pin,
                                                                      K* __restrict__ kout, uint32_t* __restrict__ pout, long long n,
                                                                      int shift, const uint32_t* __restrict__ bucket_base,
                                                                      uint32_t* __restrict__ tile_counter, uint32_t* status) {
  __shared__ uint32_t s_tile;
  __shared__ uint32_t cnt[8][256];
  __shared__ uint32_t digit_start[256];
  __shared__ uint32_t gbase[256];
  __shared__ uint32_t wsum[8];
  extern __shared__ __align__(16) uint8_t dyn[];
  K* skeys = reinterpret_cast<K*>(dyn);
  uint32_t* spay = reinterpret_cast<uint32_t*>(dyn + sizeof(K) * kSortTile);

  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  for (int i = tid; i < 8 * 256; i += kSortThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const long long base = (long long)tile * kSortTile;
  const int valid = (int)((n - base) < kSortTile ? (n - base) : kSortTile);

  K key[kSortItems];
  uint32_t pay[kSortItems];
  uint32_t rnk[kSortItems];
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const int local = w * (32 * kSortItems) + i * 32 + lane;
    if (local < valid) {
      key[i] = kin[base + local];
      pay[i] = pin[base + local];
    } else {
      key[i] = ~(K)0;  // sorts after every valid key of the tile in every pass; never written out
      pay[i] = 0;
    }
  }
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const uint32_t d = (uint32_t)(key[i] >> shift) & 0xffu;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader) {
      old = cnt[w][d];
      cnt[w][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rnk[i] = old + __popc(peers & lt_mask);
    __syncwarp();
  }
  __syncthreads();

  // thread d owns digit d: prefix over the 8 warps, publish, look back, tile-local digit offsets
  uint32_t tot = 0;
#pragma unroll
  for (int ww = 0; ww < 8; ++ww) {
    const uint32_t c = cnt[ww][tid];
    cnt[ww][tid] = tot;
    tot += c;
  }
  volatile uint32_t* vstatus = status;
  vstatus[(size_t)tile * 256 + tid] = (tile == 0 ? kFlagIncl : kFlagAgg) | tot;
  uint32_t excl = 0;
  if (tile > 0) {
    long long t = (long long)tile - 1;
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while (true) {
      const uint32_t v = vstatus[(size_t)t * 256 + tid];
      const uint32_t flag = v & ~kValMask;
      if (flag == 0) {
        if ((++spins & 0xfff) == 0) {
          if (t0 == 0) t0 = global_timer_ns();
          else if (global_timer_ns() - t0 > DFS_WAIT_LIMIT_NS) { printf("dfs_b200: radix look-back timeout\n"); __trap(); }
        }
        continue;
      }
      excl += v & kValMask;
      if (flag == kFlagIncl) break;
      --t;
    }
    vstatus[(size_t)tile * 256 + tid] = kFlagIncl | (excl + tot);
  }
  gbase[tid] = bucket_base[tid] + excl;
  // block exclusive scan of tot over the 256 digits
  uint32_t incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  uint32_t woff = 0;
#pragma unroll
  for (int ww = 0; ww < 8; ++ww) woff += (ww < w) ? wsum[ww] : 0u;
  digit_start[tid] = woff + incl - tot;
  __syncthreads();

#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const uint32_t d = (uint32_t)(key[i] >> shift) & 0xffu;
    const uint32_t pos = digit_start[d] + cnt[w][d] + rnk[i];
    skeys[pos] = key[i];
    spay[pos] = pay[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const int pos = tid + i * kSortThreads;
    if (pos < valid) {
      const K k = skeys[pos];
      const uint32_t d = (uint32_t)(k >> shift) & 0xffu;
      const size_t dst = (size_t)gbase[d] + (uint32_t)(pos - digit_start[d]);
      kout[dst] = k;
      pout[dst] = spay[pos];
    }
  }
}

// ---- FAR/FRR sweep ----------------------------------------------------------------------
struct SweepBest {
  double diff;
  long long idx;  // curve index k in [0, n]
  long long c1;   // bonafide count among the first k sorted scores
};

__global__ void __launch_bounds__(256) sweep_count_kernel(const uint32_t* __restrict__ pay, long long n, uint32_t* __restrict__ block_ones) {
  const long long base = (long long)blockIdx.x * kSortTile;
  uint32_t ones = 0;
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const long long j = base + threadIdx.x + i * 256;
    if (j < n) ones += pay[j] >> 31;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) ones += __shfl_xor_sync(0xffffffffu, ones, o);
  __shared__ uint32_t part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ones;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s += part[i];
    block_ones[blockIdx.x] = s;
  }
}

// single-block exclusive scan of the per-tile bonafide counts (<= ~260k tiles at n = 2^30)
__global__ void __launch_bounds__(1024) sweep_scan_kernel(const uint32_t* __restrict__ block_ones, long long nb,
                                                           unsigned long long* __restrict__ block_excl) {
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (long long start = 0; start < nb; start += 1024) {
    const long long i = start + threadIdx.x;
    const unsigned long long v = i < nb ? block_ones[i] : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    unsigned long long woff = 0;
    for (int ww = 0; ww < w; ++ww) woff += wsum[ww];
    if (i < nb) block_excl[i] = carry + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += woff + incl;
    __syncthreads();
  }
}

__device__ __forceinline__ bool better(double d, long long i, double bd, long long bi) { return d < bd || (d == bd && i < bi); }

__global__ void __launch_bounds__(256) sweep_min_kernel(const uint32_t* __restrict__ pay, long long n, long long n_bona, long long n_spoof,
                                                         const unsigned long long* __restrict__ block_excl, SweepBest* __restrict__ block_best) {
  // blocked arrangement: thread t owns sorted positions base + 16 t .. + 15
  const long long base = (long long)blockIdx.x * kSortTile + (long long)threadIdx.x * kSortItems;
  uint32_t lab[kSortItems];
  uint32_t ones = 0;
  if (base + kSortItems <= n) {
    const uint4* p4 = reinterpret_cast<const uint4*>(pay + base);
#pragma unroll
    for (int v = 0; v < kSortItems / 4; ++v) {
      const uint4 q = p4[v];
      lab[4 * v + 0] = q.x >> 31; lab[4 * v + 1] = q.y >> 31; lab[4 * v + 2] = q.z >> 31; lab[4 * v + 3] = q.w >> 31;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) lab[i] = (base + i < n) ? (pay[base + i] >> 31) : 0u;
  }
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) ones += lab[i];
  // block exclusive scan of `ones`
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t incl = ones;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  __shared__ uint32_t wsum[8];
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  uint32_t woff = 0;
#pragma unroll
  for (int ww = 0; ww < 8; ++ww) woff += (ww < w) ? wsum[ww] : 0u;
  long long c1 = (long long)block_excl[blockIdx.x] + woff + incl - ones;

  const double dspoof = (double)n_spoof, dbona = (double)n_bona;
  double bd = 1.0e300;
  long long bi = 0x7fffffffffffffffll, bc1 = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) { bd = 1.0; bi = 0; bc1 = 0; }  // k = 0: FAR = 1, FRR = 0
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const long long j = base + i;
    if (j < n) {
      c1 += lab[i];
      const long long k = j + 1;
      const long long c0 = k - c1;
      const double far = __ddiv_rn((double)(n_spoof - c0), dspoof);   // evaluation.py:21-23
      const double frr = __ddiv_rn((double)c1, dbona);                // evaluation.py:24-26
      const double d = fabs(__dsub_rn(far, frr));                     // evaluation.py:28
      if (better(d, k, bd, bi)) { bd = d; bi = k; bc1 = c1; }
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, bd, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const long long oc = __shfl_xor_sync(0xffffffffu, bc1, o);
    if (better(od, oi, bd, bi)) { bd = od; bi = oi; bc1 = oc; }
  }
  __shared__ SweepBest part[8];
  if (lane == 0) part[w] = SweepBest{bd, bi, bc1};
  __syncthreads();
  if (threadIdx.x == 0) {
    SweepBest b = part[0];
    for (int i = 1; i < 8; ++i)
      if (better(part[i].diff, part[i].idx, b.diff, b.idx)) b = part[i];
    block_best[blockIdx.x] = b;
  }
}

template <typename K>
__global__ void __launch_bounds__(256) sweep_final_kernel(const SweepBest* __restrict__ block_best, long long nb, const K* __restrict__ skeys,
                                                           long long n, long long n_bona, long long n_spoof, dfs_eer_result* __restrict__ res) {
  double bd = 1.0e300;
  long long bi = 0x7fffffffffffffffll, bc1 = 0;
  for (long long i = threadIdx.x; i < nb; i += blockDim.x) {
    const SweepBest b = block_best[i];
    if (better(b.diff, b.idx, bd, bi)) { bd = b.diff; bi = b.idx; bc1 = b.c1; }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, bd, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const long long oc = __shfl_xor_sync(0xffffffffu, bc1, o);
    if (better(od, oi, bd, bi)) { bd = od; bi = oi; bc1 = oc; }
  }
  __shared__ SweepBest part[8];
  if (lane == 0) part[w] = SweepBest{bd, bi, bc1};
  __syncthreads();
  if (threadIdx.x == 0) {
    SweepBest b = part[0];
    for (int i = 1; i < 8; ++i)
      if (better(part[i].diff, part[i].idx, b.diff, b.idx)) b = part[i];
    const long long k = b.idx, c0 = k - b.c1;
    const double far = __ddiv_rn((double)(n_spoof - c0), (double)n_spoof);
    const double frr = __ddiv_rn((double)b.c1, (double)n_bona);
    res->eer = __ddiv_rn(__dadd_rn(far, frr), 2.0);                  // evaluation.py:29
    typedef typename ScoreOf<K>::type S;
    const S eps = (S)1e-6;                                            // numpy: the python float is a weak scalar
    double thr;
    if (k == 0) thr = (double)(S)(from_key(skeys[0]) - eps);          // evaluation.py:32-33
    else if (k == n) thr = (double)(S)(from_key(skeys[n - 1]) + eps); // :34-35
    else thr = (double)from_key(skeys[k - 1]);                        // :37
    res->threshold = thr;
    res->eer_idx = k;
    res->n_bonafide = n_bona;
    res->n_spoof = n_spoof;
  }
}

template <typename K>
__global__ void sort_outputs_kernel(const K* __restrict__ skeys, const uint32_t* __restrict__ pay, long long n, uint32_t* __restrict__ perm,
                                    typename ScoreOf<K>::type* __restrict__ sorted) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (perm) perm[i] = pay[i] & 0x7fffffffu;
    if (sorted) sorted[i] = from_key(skeys[i]);
  }
}

// ---- workspace (grow-only, one per device) ---------------------------------------------------
struct EerWorkspace {
  void* base = nullptr;
  size_t bytes = 0;
  int device = -1;
};
static EerWorkspace g_ws[16];

static int get_workspace(size_t bytes, void** out) {
  int dev = 0;
  DFS_CUDA_CHECK(cudaGetDevice(&dev));
  DFS_REQUIRE(dev >= 0 && dev < 16, DFS_ERR_INVALID, "device index %d out of range", dev);
  EerWorkspace& w = g_ws[dev];
  if (w.bytes < bytes) {
    if (w.base) DFS_CUDA_CHECK(cudaFree(w.base));
    w.base = nullptr;
    w.bytes = 0;
    DFS_CUDA_CHECK(cudaMalloc(&w.base, bytes));
    w.bytes = bytes;
  }
  *out = w.base;
  return DFS_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename K>
static int eer_impl(const void* scores, const uint8_t* labels, int64_t n, dfs_eer_result* result_host, uint32_t* perm, void* sorted,
                    cudaStream_t stream) {
  typedef typename ScoreOf<K>::type S;
  constexpr int PASSES = sizeof(K);
  const long long tiles = ceil_div64(n, kSortTile);
  // carve the workspace
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t o_k0 = carve(sizeof(K) * n), o_k1 = carve(sizeof(K) * n);
  const size_t o_p0 = carve(4 * (size_t)n), o_p1 = carve(4 * (size_t)n);
  const size_t o_status = carve((size_t)tiles * 256 * 4);
  const size_t o_hist = carve(PASSES * 256 * 4);
  const size_t o_small = carve(256);  // [0] n_ones (u64), [8..] tile counter, [64..] skip flags
  const size_t o_bones = carve((size_t)tiles * 4), o_bexcl = carve((size_t)tiles * 8), o_bbest = carve((size_t)tiles * sizeof(SweepBest));
  const size_t o_res = carve(sizeof(dfs_eer_result));
  void* base = nullptr;
  DFS_PROPAGATE(get_workspace(off, &base));
  uint8_t* b8 = static_cast<uint8_t*>(base);
  K* keys[2] = {reinterpret_cast<K*>(b8 + o_k0), reinterpret_cast<K*>(b8 + o_k1)};
  uint32_t* pay[2] = {reinterpret_cast<uint32_t*>(b8 + o_p0), reinterpret_cast<uint32_t*>(b8 + o_p1)};
  uint32_t* status = reinterpret_cast<uint32_t*>(b8 + o_status);
  uint32_t* hist = reinterpret_cast<uint32_t*>(b8 + o_hist);
  unsigned long long* n_ones = reinterpret_cast<unsigned long long*>(b8 + o_small);
  uint32_t* tile_counter = reinterpret_cast<uint32_t*>(b8 + o_small + 8);
  uint32_t* skip = reinterpret_cast<uint32_t*>(b8 + o_small + 64);
  uint32_t* bones = reinterpret_cast<uint32_t*>(b8 + o_bones);
  unsigned long long* bexcl = reinterpret_cast<unsigned long long*>(b8 + o_bexcl);
  SweepBest* bbest = reinterpret_cast<SweepBest*>(b8 + o_bbest);
  dfs_eer_result* res_dev = reinterpret_cast<dfs_eer_result*>(b8 + o_res);

  DFS_CUDA_CHECK(cudaMemsetAsync(hist, 0, PASSES * 256 * 4, stream));
  DFS_CUDA_CHECK(cudaMemsetAsync(b8 + o_small, 0, 256, stream));
  int num_sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const unsigned prep_grid = (unsigned)std::min<long long>(ceil_div64(n, 256), (long long)num_sms * 8);
  sort_prep_kernel<K><<<prep_grid, 256, 0, stream>>>(static_cast<const S*>(scores), labels, n, keys[0], pay[0], hist, n_ones);
  DFS_LAUNCH_CHECK();
  sort_scan_hist_kernel<<<PASSES, 256, 0, stream>>>(hist, skip, n);
  DFS_LAUNCH_CHECK();
  // the label counts and skip flags decide the host control flow (single-class early-out; skipped passes)
  struct { unsigned long long ones; uint32_t counter; uint32_t pad[13]; uint32_t skip[8]; } small_host;
  DFS_CUDA_CHECK(cudaMemcpyAsync(&small_host, b8 + o_small, sizeof(small_host), cudaMemcpyDeviceToHost, stream));
  DFS_CUDA_CHECK(cudaStreamSynchronize(stream));
  const long long n_bona = (long long)small_host.ones, n_spoof = n - n_bona;

  static bool configured = false;
  const size_t dyn_smem = (sizeof(K) + 4) * kSortTile;
  if (!configured) {
    DFS_CUDA_CHECK(cudaFuncSetAttribute(onesweep_pass_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * kSortTile));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(onesweep_pass_kernel<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * kSortTile));
    configured = true;
  }
  int cur = 0;
  for (int ps = 0; ps < PASSES; ++ps) {
    if (small_host.skip[ps]) continue;  // every key shares this digit: the pass is the identity
    DFS_CUDA_CHECK(cudaMemsetAsync(status, 0, (size_t)tiles * 256 * 4, stream));
    DFS_CUDA_CHECK(cudaMemsetAsync(tile_counter, 0, 4, stream));
    onesweep_pass_kernel<K><<<(unsigned)tiles, kSortThreads, dyn_smem, stream>>>(keys[cur], pay[cur], keys[cur ^ 1], pay[cur ^ 1], n, 8 * ps,
                                                                                 hist + ps * 256, tile_counter, status);
    DFS_LAUNCH_CHECK();
    cur ^= 1;
  }
  if (perm != nullptr || sorted != nullptr) {
    sort_outputs_kernel<K><<<(unsigned)std::min<long long>(ceil_div64(n, 256), (long long)num_sms * 16), 256, 0, stream>>>(
        keys[cur], pay[cur], n, perm, static_cast<S*>(sorted));
    DFS_LAUNCH_CHECK();
  }
  if (n_bona == 0 || n_spoof == 0) {  // evaluation.py:18-19
    DFS_CUDA_CHECK(cudaStreamSynchronize(stream));
    result_host->eer = 0.0;
    result_host->threshold = 0.0;
    result_host->eer_idx = -1;
    result_host->n_bonafide = n_bona;
    result_host->n_spoof = n_spoof;
    return DFS_OK;
  }
  sweep_count_kernel<<<(unsigned)tiles, 256, 0, stream>>>(pay[cur], n, bones);
  DFS_LAUNCH_CHECK();
  sweep_scan_kernel<<<1, 1024, 0, stream>>>(bones, tiles, bexcl);
  DFS_LAUNCH_CHECK();
  sweep_min_kernel<<<(unsigned)tiles, 256, 0, stream>>>(pay[cur], n, n_bona, n_spoof, bexcl, bbest);
  DFS_LAUNCH_CHECK();
  sweep_final_kernel<K><<<1, 256, 0, stream>>>(bbest, tiles, keys[cur], n, n_bona, n_spoof, res_dev);
  DFS_LAUNCH_CHECK();
  DFS_CUDA_CHECK(cudaMemcpyAsync(result_host, res_dev, sizeof(dfs_eer_result), cudaMemcpyDeviceToHost, stream));
  DFS_CUDA_CHECK(cudaStreamSynchronize(stream));
  return DFS_OK;
}

int eer_device(const void* scores, int key_bytes, const uint8_t* labels, int64_t n, dfs_eer_result* result_host, uint32_t* perm,
               void* sorted, cudaStream_t stream) {
  DFS_REQUIRE(scores && labels && result_host, DFS_ERR_INVALID, "dfs_eer: NULL argument");
  DFS_REQUIRE(n > 0 && n < (1ll << 30), DFS_ERR_INVALID, "dfs_eer: n = %lld outside [1, 2^30)", (long long)n);
  DFS_REQUIRE(key_bytes == 4 || key_bytes == 8, DFS_ERR_INVALID, "dfs_eer: key_bytes must be 4 (fp32) or 8 (fp64)");
  return key_bytes == 4 ? eer_impl<uint32_t>(scores, labels, n, result_host, perm, sorted, stream)
                        : eer_impl<uint64_t>(scores, labels, n, result_host, perm, sorted, stream);
}

// ---- confusion counts -----------------------------------------------------------------------
template <typename S>
__global__ void __launch_bounds__(256) confusion_kernel(const S* __restrict__ scores, const uint8_t* __restrict__ labels, long long n, S thr,
                                                         unsigned long long* __restrict__ out4) {
  uint32_t c[4] = {0, 0, 0, 0};  // tp fp tn fn
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const bool pred = scores[i] > thr;       // evaluation.py:46
    const bool pos = labels[i] == 1;
    const bool neg = labels[i] == 0;
    c[0] += pred && pos;
    c[1] += pred && neg;
    c[2] += !pred && neg;
    c[3] += !pred && pos;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) c[k] += __shfl_xor_sync(0xffffffffu, c[k], o);
    if ((threadIdx.x & 31) == 0 && c[k]) atomicAdd(&out4[k], (unsigned long long)c[k]);
  }
}

int confusion_device(const void* scores, int key_bytes, const uint8_t* labels, int64_t n, double thr, int64_t* out4_host, cudaStream_t stream) {
  DFS_REQUIRE(scores && labels && out4_host && n >= 0, DFS_ERR_INVALID, "dfs_confusion: bad argument");
  DFS_REQUIRE(key_bytes == 4 || key_bytes == 8, DFS_ERR_INVALID, "dfs_confusion: key_bytes must be 4 or 8");
  void* base = nullptr;
  DFS_PROPAGATE(get_workspace(256, &base));
  unsigned long long* d4 = static_cast<unsigned long long*>(base);
  DFS_CUDA_CHECK(cudaMemsetAsync(d4, 0, 32, stream));
  if (n > 0) {
    const unsigned grid = (unsigned)std::min<long long>(ceil_div64(n, 256), 148 * 16);
    if (key_bytes == 4)
      confusion_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(scores), labels, n, (float)thr, d4);
    else
      confusion_kernel<double><<<grid, 256, 0, stream>>>(static_cast<const double*>(scores), labels, n, thr, d4);
    DFS_LAUNCH_CHECK();
  }
  unsigned long long h4[4];
  DFS_CUDA_CHECK(cudaMemcpyAsync(h4, d4, 32, cudaMemcpyDeviceToHost, stream));
  DFS_CUDA_CHECK(cudaStreamSynchronize(stream));
  for (int k = 0; k < 4; ++k) out4_host[k] = (int64_t)h4[k];
  return DFS_OK;
}

// ---- blend (float64) ------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ordered_bits(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double from_ordered_bits(unsigned long long k) {
  return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}

__global__ void __launch_bounds__(256) minmax_kernel(const double* __restrict__ x, long long n, unsigned long long* __restrict__ mm /*[min,max]*/) {
  unsigned long long lo = ~0ull, hi = 0ull;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long k = ordered_bits(x[i]);
    lo = k < lo ? k : lo;
    hi = k > hi ? k : hi;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const unsigned long long ol = __shfl_xor_sync(0xffffffffu, lo, o), oh = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = ol < lo ? ol : lo;
    hi = oh > hi ? oh : hi;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&mm[0], lo);
    atomicMax(&mm[1], hi);
  }
}

struct BlendParams {
  const double* src[8];
  double weight[8];
  int minmax[8];
  int m;
  double divisor;
};

__global__ void __launch_bounds__(256) blend_kernel(const __grid_constant__ BlendParams p, const unsigned long long* __restrict__ mm, long long n,
                                                    double* __restrict__ out) {
  double lo[8], range[8];
  bool flat[8];
  for (int k = 0; k < p.m; ++k) {
    lo[k] = 0.0; range[k] = 1.0; flat[k] = false;
    if (p.minmax[k]) {
      lo[k] = from_ordered_bits(mm[2 * k]);
      const double hi = from_ordered_bits(mm[2 * k + 1]);
      range[k] = __dsub_rn(hi, lo[k]);
      flat[k] = range[k] < 1e-12;                                     // predict_hybrid.py:83-84
    }
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < p.m; ++k) {
      double v = p.src[k][i];
      if (p.minmax[k]) v = flat[k] ? 0.0 : __ddiv_rn(__dsub_rn(v, lo[k]), range[k]);   // :85
      const double term = __dmul_rn(p.weight[k], v);
      acc = (k == 0) ? term : __dadd_rn(acc, term);                   // :151 / ensemble.py:121
    }
    out[i] = __ddiv_rn(acc, p.divisor);
  }
}

int blend_device(const double* const* scores, int m, const double* weights, const int* minmax, double divisor, int64_t n, double* out,
                 cudaStream_t stream) {
  DFS_REQUIRE(scores && weights && minmax && out, DFS_ERR_INVALID, "dfs_blend: NULL argument");
  DFS_REQUIRE(m >= 1 && m <= 8, DFS_ERR_INVALID, "dfs_blend: m = %d outside [1, 8]", m);
  DFS_REQUIRE(n >= 0, DFS_ERR_INVALID, "dfs_blend: negative n");
  if (n == 0) return DFS_OK;
  void* base = nullptr;
  DFS_PROPAGATE(get_workspace(256, &base));
  unsigned long long* mm = static_cast<unsigned long long*>(base);
  unsigned long long init[16];
  for (int k = 0; k < 8; ++k) { init[2 * k] = ~0ull; init[2 * k + 1] = 0ull; }
  DFS_CUDA_CHECK(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, stream));
  DFS_CUDA_CHECK(cudaStreamSynchronize(stream));  // `init` is a stack buffer
  BlendParams p{};
  p.m = m;
  p.divisor = divisor;
  const unsigned grid = (unsigned)std::min<long long>(ceil_div64(n, 256), 148 * 16);
  for (int k = 0; k < m; ++k) {
    DFS_REQUIRE(scores[k] != nullptr, DFS_ERR_INVALID, "dfs_blend: scores[%d] is NULL", k);
    p.src[k] = scores[k];
    p.weight[k] = weights[k];
    p.minmax[k] = minmax[k];
    if (minmax[k]) {
      minmax_kernel<<<grid, 256, 0, stream>>>(scores[k], n, mm + 2 * k);
      DFS_LAUNCH_CHECK();
    }
  }
  blend_kernel<<<grid, 256, 0, stream>>>(p, mm, n, out);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

__global__ void widen_kernel(const float* __restrict__ in, long long n, double* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = (double)in[i];
}

int widen_device(const float* in, int64_t n, double* out, cudaStream_t stream) {
  DFS_REQUIRE(in && out && n >= 0, DFS_ERR_INVALID, "dfs_widen: bad argument");
  if (n == 0) return DFS_OK;
  widen_kernel<<<(unsigned)std::min<long long>(ceil_div64(n, 256), 148 * 16), 256, 0, stream>>>(in, n, out);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}

}  // namespace dfs
