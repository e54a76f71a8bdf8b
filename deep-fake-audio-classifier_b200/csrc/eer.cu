// eer.cu -- the metric half of the hot path on the device:
//   calculate_eer            /root/reference/scripts/evaluation.py:7-39 (== src/evaluation.py:12-48)
//   confusion_at_threshold   /root/reference/scripts/evaluation.py:42-56
//   normalise_01 + blend     /root/reference/src/predict_hybrid.py:81-85,149-151, src/ensemble.py:121
//
// EER = stable LSD radix sort (8-bit digits, key = order-preserving integer image of the fp32/fp64
// score, payload = original index | label<<31) followed by a prefix-count FAR/FRR sweep in IEEE fp64
// and a (value, lowest index) arg-min.  A radix pass is ONE scatter kernel (radix_onesweep_kernel:
// tiles ticketed in input order, decoupled look-back, the next pass's histogram taken on the way);
// round 1's count -> scan -> scatter kernels over per-CTA super-tiles stay as a cross-check form.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace dfs {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096

// ---- order-preserving key transforms (-0.0 is canonicalised to +0.0: numpy treats them as ties) ----
// NaN of either sign maps to the largest key: np.argsort (scripts/evaluation.py:11) sorts NaN after +inf, whatever its sign
// bit or payload; from_key() of that key is a (canonical) NaN again.
__device__ __forceinline__ uint32_t to_key(float s) {
  uint32_t b = __float_as_uint(s);
  if ((b & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;
  if (b == 0x80000000u) b = 0;
  return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);   // negative: flip all bits, else set the sign bit
}
__device__ __forceinline__ uint64_t to_key(double s) {
  uint64_t b = (uint64_t)__double_as_longlong(s);
  if ((b & 0x7fffffffffffffffull) > 0x7ff0000000000000ull) return ~0ull;
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

// ---- pass 0: keys, payloads, label count, AND / OR of all keys --------------------------------
// A radix pass whose digit is identical in every key is the identity; that is the case exactly when the bitwise AND
// and OR of all keys agree on that byte (common: sigmoid scores share their exponent byte), so the host can skip it.
struct SortHeader {
  unsigned long long ones;      // number of labels != 0
  unsigned long long key_and;   // AND of all keys (zero-extended)
  unsigned long long key_or;    // OR of all keys
};

template <typename K>
__global__ void __launch_bounds__(256) sort_prep_kernel(const typename ScoreOf<K>::type* __restrict__ scores,
                                                         const uint8_t* __restrict__ labels, long long n, K* __restrict__ keys,
                                                         uint32_t* __restrict__ pay, SortHeader* __restrict__ hdr) {
  uint32_t ones = 0;
  K kand = ~(K)0, kor = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const K k = to_key(scores[i]);
    const uint32_t lab = labels[i] != 0;
    keys[i] = k;
    pay[i] = (uint32_t)i | (lab << 31);
    ones += lab;
    kand &= k;
    kor |= k;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    ones += __shfl_xor_sync(0xffffffffu, ones, o);
    kand &= (K)__shfl_xor_sync(0xffffffffu, (unsigned long long)kand, o);
    kor |= (K)__shfl_xor_sync(0xffffffffu, (unsigned long long)kor, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (ones) atomicAdd(&hdr->ones, (unsigned long long)ones);
    atomicAnd(&hdr->key_and, (unsigned long long)kand);
    atomicOr(&hdr->key_or, (unsigned long long)kor);
  }
}

// ---- one radix pass = upsweep (per-super-tile digit counts) -> scan -> downsweep (stable scatter) ----
// A super-tile is kSuperTiles consecutive tiles handled by ONE CTA in order, carrying its per-digit global write cursors in
// shared memory, so no CTA ever waits on another one (a decoupled look-back walks hundreds of predecessors when ~600 tiles
// are in flight: measured 8.5 ms per 100 M keys; this form is bandwidth-bound).
constexpr int kSuperTiles = 16;
constexpr int kSuperKeys = kSuperTiles * kSortTile;   // 65,536 keys per CTA

// lanes holding the same 8-bit digit (8 ballots; no shared-memory atomics: those cost ~2 cycles per active lane)
__device__ __forceinline__ uint32_t warp_peers8(uint32_t d, bool ok) {
  uint32_t peers = __ballot_sync(0xffffffffu, ok);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    // bal = lanes whose bit b is set; m = 0 if mine is set, ~0 if not: peers &= (mine set ? bal : ~bal).  Written with the predicate kept
    // in the asm block so that ptxas moves the digit's bits into predicates with one R2P and spends VOTE + predicated NOT + AND per bit
    // (29 instructions per 32 keys; the C++ forms `bit ? bal : ~bal` and `~(bal ^ -bit)` compile to 51)
    uint32_t bal, m;
    asm volatile("{\n .reg .pred p;\n .reg .b32 t;\n and.b32 t, %2, %3;\n setp.ne.b32 p, t, 0;\n vote.sync.ballot.b32 %0, p, 0xffffffff;\n"
                 " selp.b32 %1, 0, 0xffffffff, p;\n}" : "=r"(bal), "=r"(m) : "r"(d), "r"(1u << b));
    peers &= bal ^ m;
  }
  return peers;
}

// ---- digit histograms with thread-private byte counters ---------------------------------------------------
// Ballot ranking costs ~70 warp instructions per 32 keys and shared-memory atomics ~2 cycles per lane; a histogram needs
// neither: thread t owns the byte column t of a [bins][kCntStride] table, so an increment is an unsynchronised
// LDS.U8 / IADD / STS.U8 (rows 260 B apart: a warp's 32 accesses spread over the banks like 32 random words).  A thread
// may add at most 255 keys between two flushes; a flush sums each row with dp4a (rows are conflict-free: word j of row r
// sits in bank (r + j) % 32) and clears it.
constexpr int kCntStride = 260;

template <typename K> struct VecOf;
template <> struct VecOf<uint32_t> { static constexpr int V = 4; };
template <> struct VecOf<uint64_t> { static constexpr int V = 2; };

template <typename T, int V>
__device__ __forceinline__ void load_vec(const T* __restrict__ p, T (&out)[V]) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  if constexpr (sizeof(T) == 4) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) out[e] = *reinterpret_cast<const T*>(&w[e]);
  } else {
    const uint64_t w[2] = {((uint64_t)q.y << 32) | q.x, ((uint64_t)q.w << 32) | q.z};
#pragma unroll
    for (int e = 0; e < 2; ++e) out[e] = *reinterpret_cast<const T*>(&w[e]);
  }
}

// sum and clear row `row` (256 counters) of a private-counter table
__device__ __forceinline__ uint32_t flush_counter_row(uint8_t* cnt, int row) {
  uint32_t* w = reinterpret_cast<uint32_t*>(cnt + row * kCntStride);
  uint32_t tot = 0;
#pragma unroll 16
  for (int j = 0; j < 64; ++j) {
    tot = __dp4a(w[j], 0x01010101u, tot);
    w[j] = 0u;
  }
  return tot;
}

template <typename K>
__global__ void __launch_bounds__(256) radix_upsweep_kernel(const K* __restrict__ kin, long long n, int shift, uint32_t* __restrict__ counts /*[supers][256]*/) {
  extern __shared__ __align__(16) uint8_t cnt[];   // [256][kCntStride]
  constexpr int V = VecOf<K>::V;
  constexpr int kFlushKeys = 128;                  // keys per thread between flushes
  for (int i = threadIdx.x; i < 256 * kCntStride / 4; i += 256) reinterpret_cast<uint32_t*>(cnt)[i] = 0u;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kSuperKeys;
  const long long end = (base + kSuperKeys) < n ? (base + kSuperKeys) : n;
  uint8_t* mine = cnt + threadIdx.x;
  uint32_t tot = 0;
  for (long long c0 = base; c0 < end; c0 += 256ll * kFlushKeys) {
#pragma unroll 4
    for (int it = 0; it < kFlushKeys / V; ++it) {
      const long long i = c0 + ((long long)it * 256 + threadIdx.x) * V;
      if (i + V <= end) {
        K k[V];
        load_vec<K, V>(kin + i, k);
#pragma unroll
        for (int e = 0; e < V; ++e) mine[((uint32_t)(k[e] >> shift) & 0xffu) * kCntStride] += 1;
      } else {
        for (long long j = i; j < end; ++j) mine[((uint32_t)(kin[j] >> shift) & 0xffu) * kCntStride] += 1;
      }
    }
    __syncthreads();
    tot += flush_counter_row(cnt, threadIdx.x);
    __syncthreads();
  }
  counts[(size_t)blockIdx.x * 256 + threadIdx.x] = tot;
}

// counts[s][d] -> exclusive prefix over super-tiles s (per digit), in place; bucket_base[d] = exclusive prefix over digits of the
// per-digit totals.  One block per digit: thread t owns the consecutive super-tiles [t * per, (t + 1) * per) (loads issued together,
// then a block scan of the 256 partial sums), the last block to finish scans the 256 digit totals.  The first version walked the
// 1,526 rows of a 100 M-key sort in ONE block, one dependent L2 round trip per row: ~0.1-0.2 ms per pass for 1.5 MB of counters.
__global__ void __launch_bounds__(256) radix_scan_kernel(uint32_t* __restrict__ counts, int supers, uint32_t* __restrict__ bucket_base,
                                                          uint32_t* __restrict__ digit_total /*[256]*/, unsigned int* __restrict__ done_counter) {
  constexpr int kMaxPer = 64;                       // 256 threads x 64 super-tiles x 65,536 keys = 2^30 keys
  const int d = blockIdx.x, t = threadIdx.x;
  const int per = (supers + 255) / 256;
  const int s0 = t * per, s1 = (s0 + per) < supers ? (s0 + per) : supers;
  uint32_t v[kMaxPer];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < kMaxPer; ++k) {
    v[k] = 0;
    if (k < per && s0 + k < s1) v[k] = counts[(size_t)(s0 + k) * 256 + d];
  }
#pragma unroll
  for (int k = 0; k < kMaxPer; ++k) sum += v[k];
  __shared__ uint32_t wsum[8];
  __shared__ bool last;
  const int lane = t & 31, w = t >> 5;
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  uint32_t woff = 0, total = 0;
#pragma unroll
  for (int ww = 0; ww < 8; ++ww) {
    woff += (ww < w) ? wsum[ww] : 0u;
    total += wsum[ww];
  }
  uint32_t run = woff + incl - sum;
#pragma unroll
  for (int k = 0; k < kMaxPer; ++k) {
    if (k < per && s0 + k < s1) {
      counts[(size_t)(s0 + k) * 256 + d] = run;
      run += v[k];
    }
  }
  if (t == 0) {
    digit_total[d] = total;
    __threadfence();
    last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  const uint32_t mine = *reinterpret_cast<volatile uint32_t*>(digit_total + t);
  uint32_t inc2 = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, inc2, o);
    if (lane >= o) inc2 += up;
  }
  __syncthreads();
  if (lane == 31) wsum[w] = inc2;
  __syncthreads();
  uint32_t woff2 = 0;
#pragma unroll
  for (int ww = 0; ww < 8; ++ww) woff2 += (ww < w) ? wsum[ww] : 0u;
  bucket_base[t] = woff2 + inc2 - mine;
  if (t == 0) *done_counter = 0;                    // self-resetting for the next pass
}

template <typename K>
__global__ void __launch_bounds__(kSortThreads, 4) radix_downsweep_kernel(const K* __restrict__ kin, const uint32_t* __restrict__ pin,
                                                                        K* __restrict__ kout, uint32_t* __restrict__ pout, long long n,
                                                                        int shift, const uint32_t* __restrict__ bucket_base,
                                                                        const uint32_t* __restrict__ super_prefix) {
  __shared__ uint32_t cnt[8][256];
  __shared__ uint32_t digit_start[256];
  __shared__ uint32_t gbase[256];
  __shared__ uint32_t cursor[256];   // next global write position of each digit for this super-tile
  __shared__ uint32_t wsum[8];
  extern __shared__ __align__(16) uint8_t dyn[];
  K* skeys = reinterpret_cast<K*>(dyn);
  uint32_t* spay = reinterpret_cast<uint32_t*>(dyn + sizeof(K) * kSortTile);

  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  cursor[tid] = bucket_base[tid] + super_prefix[(size_t)blockIdx.x * 256 + tid];
  const uint32_t lt_mask = (1u << lane) - 1u;

  for (int sub = 0; sub < kSuperTiles; ++sub) {
    const long long base = (long long)blockIdx.x * kSuperKeys + (long long)sub * kSortTile;
    if (base >= n) break;
    const int valid = (int)((n - base) < kSortTile ? (n - base) : kSortTile);
    for (int i = tid; i < 8 * 256; i += kSortThreads) (&cnt[0][0])[i] = 0;
    // L2 prefetch of the NEXT tile's keys and payloads (two 32-byte sectors of each per thread): ncu showed the warps 35 % of their time on
    // the first use of the key loads and 12 % on the payload loads -- the pass is latency-bound at 37 % of the copy bandwidth
    if (sub + 1 < kSuperTiles) {
      const long long nb = base + kSortTile;
      if (nb < n) {
        const long long lim = n - nb < kSortTile ? n - nb : kSortTile;          // keys of the next tile
        constexpr int KJ = 2 * (int)sizeof(K) / 4;                                // key sectors per thread: 16 / 32 KB per tile
#pragma unroll
        for (int j = 0; j < KJ; ++j) {
          const long long e = (long long)(KJ * tid + j) * (32 / (int)sizeof(K));  // first element of this sector
          if (e < lim) asm volatile("prefetch.global.L2 [%0];" ::"l"(kin + nb + e) : "memory");
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const long long ep = (long long)(2 * tid + j) * 8;
          if (ep < lim) asm volatile("prefetch.global.L2 [%0];" ::"l"(pin + nb + ep) : "memory");
        }
      }
    }
    __syncthreads();

    // keys only: the payloads are fetched after the ranking, straight into their staging slot, so they are not live in
    // registers across the ballot loop (<= 64 registers -> 4 CTAs per SM instead of 2)
    K key[kSortItems];
    uint32_t rnk[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const int local = w * (32 * kSortItems) + i * 32 + lane;
      key[i] = (local < valid) ? kin[base + local] : ~(K)0;  // padding sorts after every valid key of the tile; never written out
    }
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t d = (uint32_t)(key[i] >> shift) & 0xffu;
      const uint32_t peers = warp_peers8(d, true);
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

    // thread d owns digit d: prefix over the 8 warps, tile-local digit offsets, global cursor
    uint32_t tot = 0;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) {
      const uint32_t c = cnt[ww][tid];
      cnt[ww][tid] = tot;
      tot += c;
    }
    // padding keys (digit 255, ranked last) are counted in tot but never written; keep them out of the cursor
    const uint32_t pad = (tid == 255) ? (uint32_t)(kSortTile - valid) : 0u;
    gbase[tid] = cursor[tid];
    cursor[tid] += tot - pad;
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
      const int local = w * (32 * kSortItems) + i * 32 + lane;
      skeys[pos] = key[i];
      spay[pos] = (local < valid) ? pin[base + local] : 0u;
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
    __syncthreads();
  }
}

// ---- the same LSD sort as ONE scatter kernel per pass ("onesweep") ------------------------------------------------------------
// What the super-tile form above costs at 100 M keys (profiles/r02v): every CTA owns 256 private output streams, so ~600 resident CTAs
// write 150 k streams in 64-byte pieces -- partial 32-byte sectors that leave the L2 before the same CTA's next tile completes them
// (ECC DRAM: read-modify-write, +40 % traffic) and no DRAM page locality -- plus a counting pass per digit (0.17 ms each).
// Here the tiles are handed out in INPUT ORDER (atomic ticket), so all resident CTAs work on one window of consecutive tiles and their
// runs of a digit are adjacent in the output and written within microseconds of each other; a tile's global offset per digit comes from a
// decoupled look-back over the tiles before it (status word = 2 flag bits + 30-bit count, hence n < 2^30; a ticket holder is always a
// running CTA, so the chain cannot dead-lock; the first wave resolves in ~sqrt(#resident) steps, then the chain is 1-2 tiles deep).
// No counting pass: the global histogram of the NEXT pass's digit is taken by the same kernel from the ballots it already issues
// (per-warp counters kept across the CTA's tiles, 256 atomics per CTA at the end); the first pass's histogram comes from the prep
// kernel.  Tile = 512 threads x 16 keys = 8,192 (128-byte runs per digit); 2 CTAs per SM.
constexpr uint32_t kOsAgg = 1u << 30, kOsIncl = 1u << 31, kOsMask = kOsAgg - 1u;
template <typename K, int T> struct OsCfg {
  static constexpr int I = sizeof(K) == 4 ? 16 : 8;   // keys per thread
  static constexpr int TILE = T * I;
  static constexpr int W = T / 32;
  static constexpr int CTAS = 1024 / T;               // per SM: 64 registers per thread
  static constexpr size_t SMEM = (sizeof(K) + 4) * (size_t)TILE;
};

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// exclusive scan of one value per thread over the first 256 threads of the block (8 warps); `wsum` = 8 words of shared memory.
// Called by ALL threads of the block (it contains a barrier); threads >= 256 pass 0 and ignore the result.
__device__ __forceinline__ uint32_t block_excl_scan256(uint32_t v, uint32_t* wsum) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31 && w < 8) wsum[w] = incl;
  __syncthreads();
  uint32_t woff = 0;
#pragma unroll
  for (int ww = 0; ww < 8; ++ww) woff += (ww < w) ? wsum[ww] : 0u;
  return woff + incl - v;
}

// Schedule of a CTA: the look-back of a tile sits behind the RANKING OF THE CTA'S NEXT TILE -- stage tile t, rank tile t' (publishing its
// aggregate), then look back for t and write it out.  A tile's aggregate is thus public ~0.8 of a tile time before the tiles after it
// ask for it (with the look-back right after a tile's own staging, ncu showed 35 % of the warp samples spinning in it: ~12 hops and
// ~40 polls per tile), at no cost in shared memory: the per-warp counter rows are private to their warp between barriers, and the key
// registers of t are free once staged (the keys of t' are loaded item by item behind the staging stores).
// HIST: how the next pass's histogram is taken -- 0: not here (radix_hist_kernel runs before each pass), 1: one shared-memory atomic per key,
// 2: ballot peers + per-warp counters (no atomics, ~45 more instructions per 32 keys).
// T: threads per CTA (512: tiles of 8,192 fp32 keys, 2 CTAs per SM; 256: 4,096 keys, 4 CTAs per SM).
// FIRST: the pass reads the caller's scores and labels instead of a key / payload image (key = to_key(score), payload = index | label << 31),
// so no such image is written and read back before the first pass (8 + 8 B per score).  The score bits are converted to keys when the
// ranking starts, not where they are loaded: converting in the load loop made every load wait for its data (0.84 ms per pass).
template <typename K, int HIST, int T, bool FIRST>
__global__ void __launch_bounds__(T, 1024 / T) radix_onesweep_kernel(const K* __restrict__ kin, const uint32_t* __restrict__ pin, K* __restrict__ kout,
                                                                     uint32_t* __restrict__ pout, uint32_t n /* < 2^30 */, int shift, int next_shift,
                                                                     const uint32_t* __restrict__ hist_cur, uint32_t* __restrict__ hist_next,
                                                                     uint32_t* __restrict__ status /*[tiles][256]*/, unsigned int* __restrict__ ticket) {
  typedef typename ScoreOf<K>::type S;
  auto load_key = [&](uint32_t idx) -> K { return kin[idx]; };   // FIRST: kin = the scores' bits
  auto load_pay = [&](uint32_t idx) -> uint32_t {
    if constexpr (FIRST) return idx | ((uint32_t)(reinterpret_cast<const uint8_t*>(pin)[idx] != 0) << 31);   // pin = the labels
    else return pin[idx];
  };
  constexpr int I = OsCfg<K, T>::I, TILE = OsCfg<K, T>::TILE, W = OsCfg<K, T>::W;
  __shared__ uint32_t cnt[W][256];
  __shared__ uint32_t nxt[HIST == 2 ? W : 1][256];
  __shared__ uint32_t digit_start[256];
  __shared__ uint32_t gbase[256];
  __shared__ uint32_t bbase[256];
  __shared__ uint32_t wsum[8];
  __shared__ uint32_t s_tile[2];
  extern __shared__ __align__(16) uint8_t dyn[];
  K* skeys = reinterpret_cast<K*>(dyn);
  uint32_t* spay = reinterpret_cast<uint32_t*>(dyn + sizeof(K) * TILE);

  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t tiles = (n + TILE - 1) / TILE;
  const int mine0 = w * (32 * I) + lane;
  // key byte -> digit with one PRMT (selector 0x444b: byte b, zero-extended) instead of shift + mask; 64-bit keys pick their half first
  const uint32_t sel = 0x4440u | (((uint32_t)shift & 31u) >> 3), sel2 = 0x4440u | (((uint32_t)next_shift & 31u) >> 3);
  auto digit = [&](K k2, uint32_t selector, int sh) -> uint32_t {
    const uint32_t word = sizeof(K) == 8 ? (uint32_t)((unsigned long long)k2 >> (sh & 32)) : (uint32_t)k2;
    return __byte_perm(word, 0u, selector);
  };

  for (int i = tid; i < (HIST == 2 ? W : 1) * 256; i += T) (&nxt[0][0])[i] = 0u;
  for (int i = tid; i < W * 256; i += T) (&cnt[0][0])[i] = 0u;
  {
    const uint32_t tot = tid < 256 ? hist_cur[tid] : 0u;
    const uint32_t ex = block_excl_scan256(tot, wsum);
    if (tid < 256) bbase[tid] = ex;
  }
  if (tid == 0) {
    s_tile[0] = atomicAdd(ticket, 1u);
    s_tile[1] = atomicAdd(ticket, 1u);
  }
  __syncthreads();

  K key[I];
  uint32_t rnk2[I / 2];
  // ranks of key[] within this warp's digit counts (cnt row w, zero on entry); next pass's histogram on the way
  auto rank_tile = [&](int valid) {
    if constexpr (FIRST) {
#pragma unroll
      for (int i = 0; i < I; ++i)
        if (mine0 + 32 * i < valid) {
          S sc;
          memcpy(&sc, &key[i], sizeof(K));
          key[i] = to_key(sc);
        }
    }
#pragma unroll
    for (int i = 0; i < I; ++i) {
      const uint32_t d = digit(key[i], sel, shift);
      const uint32_t peers = warp_peers8(d, true);
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (lane == leader) {
        old = cnt[w][d];
        cnt[w][d] = old + __popc(peers);
      }
      old = __shfl_sync(0xffffffffu, old, leader);
      const uint32_t r = old + __popc(peers & lt_mask);
      rnk2[i >> 1] = (i & 1) ? (rnk2[i >> 1] | (r << 16)) : r;
      if (HIST != 0 && next_shift >= 0) {   // block-uniform
        const bool ok = (mine0 + 32 * i) < valid;
        const uint32_t d2 = digit(key[i], sel2, next_shift);
        if constexpr (HIST == 2) {
          const uint32_t p2 = warp_peers8(d2, ok);
          if (ok && lane == __ffs(p2) - 1) nxt[w][d2] += __popc(p2);
        } else {
          if (ok) atomicAdd(&nxt[0][d2], 1u);
        }
      }
      __syncwarp();
    }
  };
  // thread d < 256: exclusive prefix of digit d over the warps (left in cnt), the tile's count published; returns the padded total
  auto publish_tile = [&](uint32_t tile, int valid, uint32_t& count_d) -> uint32_t {
    uint32_t tot = 0;
#pragma unroll
    for (int ww = 0; ww < W; ++ww) {
      const uint32_t c = cnt[ww][tid];
      cnt[ww][tid] = tot;
      tot += c;
    }
    count_d = tot - ((tid == 255) ? (uint32_t)(TILE - valid) : 0u);
    st_relaxed_u32(status + (size_t)tile * 256 + tid, (tile == 0 ? kOsIncl : kOsAgg) | count_d);
    return tot;
  };

  uint32_t tile = s_tile[0];
  uint32_t count_d = 0;
  if (tile < tiles) {   // block-uniform
    {
      const uint32_t base = tile * TILE;
      const int valid = (int)((n - base) < TILE ? (n - base) : TILE);
#pragma unroll
      for (int i = 0; i < I; ++i) key[i] = (mine0 + 32 * i < valid) ? load_key(base + mine0 + 32 * i) : ~(K)0;
      rank_tile(valid);
      __syncthreads();
      uint32_t tot = 0;
      if (tid < 256) tot = publish_tile(tile, valid, count_d);
      const uint32_t ex = block_excl_scan256(tot, wsum);
      if (tid < 256) digit_start[tid] = ex;
      __syncthreads();
    }
    int buf = 0;
    while (true) {
      // here: key / rnk2 / cnt / digit_start / count_d belong to `tile`, whose aggregate is public
      const uint32_t ntile = s_tile[buf ^ 1];
      const bool has_next = ntile < tiles;
      const uint32_t base = tile * TILE, nbase = ntile * TILE;
      const int valid = (int)((n - base) < TILE ? (n - base) : TILE);
      const int nvalid = has_next ? (int)((n - nbase) < TILE ? (n - nbase) : TILE) : 0;
      if (tid == 0) s_tile[buf] = has_next ? atomicAdd(ticket, 1u) : 0xffffffffu;   // the ticket after ntile (slot of `tile`: read a loop ago)

      // stage `tile`; each key register is refilled with the next tile's key as soon as it is stored
#pragma unroll
      for (int i = 0; i < I; ++i) {
        const uint32_t d = digit(key[i], sel, shift);
        const uint32_t pos = digit_start[d] + cnt[w][d] + ((i & 1) ? (rnk2[i >> 1] >> 16) : (rnk2[i >> 1] & 0xffffu));
        const uint32_t sw = pos ^ ((pos >> 5) & 31u);
        skeys[sw] = key[i];
        spay[sw] = (mine0 + 32 * i < valid) ? load_pay(base + mine0 + 32 * i) : 0u;
        if (has_next) key[i] = (mine0 + 32 * i < nvalid) ? load_key(nbase + mine0 + 32 * i) : ~(K)0;
      }
      if (has_next) {   // the next tile's payloads (labels) into the L2
        if constexpr (FIRST) {
          if (tid < TILE / 32 && tid * 32 < nvalid) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint8_t*>(pin) + nbase + tid * 32) : "memory");
        } else {
          constexpr int PJ = TILE * 4 / 32 / T;
#pragma unroll
          for (int j = 0; j < PJ; ++j) {
            const uint32_t e = (uint32_t)(PJ * tid + j) * 8;
            if ((int)e < nvalid) asm volatile("prefetch.global.L2 [%0];" ::"l"(pin + nbase + e) : "memory");
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) cnt[w][lane + 32 * j] = 0u;   // this warp's row: nobody else touches it before the next barrier
      __syncwarp();
      if (has_next) rank_tile(nvalid);
      __syncthreads();   // `tile` staged, `ntile` ranked

      uint32_t tot = 0, ncount = 0;
      if (tid < 256) {
        if (has_next) tot = publish_tile(ntile, nvalid, ncount);
        uint32_t excl = 0;
        if (tile > 0) {
          const uint32_t* sp = status + (size_t)(tile - 1) * 256 + tid;
          while (true) {
            uint32_t v = ld_relaxed_u32(sp);
            while ((v >> 30) == 0u) {
              __nanosleep(40);
              v = ld_relaxed_u32(sp);
            }
            excl += v & kOsMask;
            if (v & kOsIncl) break;
            sp -= 256;
          }
          st_relaxed_u32(status + (size_t)tile * 256 + tid, kOsIncl | (excl + count_d));
        }
        gbase[tid] = bbase[tid] + excl - digit_start[tid];
      }
      {
        const uint32_t ex = block_excl_scan256(tot, wsum);   // barrier inside: every thread of the first 256 has read its digit_start
        if (tid < 256) digit_start[tid] = ex;
      }
      __syncthreads();   // gbase of `tile`, digit_start of `ntile`

      if (valid == TILE) {   // block-uniform; all but the last tile: no bounds predicates, the swizzled slot from G precomputed lane offsets
        constexpr int G = 32 / W;   // (pos >> 5) & 31 = (w + i W) & 31 takes G values
        uint32_t lx[G];
#pragma unroll
        for (int j = 0; j < G; ++j) lx[j] = 32u * w + ((uint32_t)lane ^ ((uint32_t)(w + j * W) & 31u));
#pragma unroll
        for (int i = 0; i < I; ++i) {
          const uint32_t sw = lx[i % G] + (uint32_t)(i * T);
          const K k = skeys[sw];
          const uint32_t dst = gbase[digit(k, sel, shift)] + (uint32_t)(tid + i * T);
          kout[dst] = k;
          pout[dst] = spay[sw];
        }
      } else {
#pragma unroll
        for (int i = 0; i < I; ++i) {
          const int pos = tid + i * T;
          if (pos < valid) {
            const uint32_t sw = (uint32_t)pos ^ (((uint32_t)pos >> 5) & 31u);
            const K k = skeys[sw];
            const uint32_t dst = gbase[digit(k, sel, shift)] + (uint32_t)pos;
            kout[dst] = k;
            pout[dst] = spay[sw];
          }
        }
      }
      __syncthreads();   // staging buffer free
      if (!has_next) break;
      tile = ntile;
      count_d = ncount;
      buf ^= 1;
    }
  }

  if (HIST != 0 && next_shift >= 0 && tid < 256) {
    uint32_t s = 0;
#pragma unroll
    for (int ww = 0; ww < (HIST == 2 ? W : 1); ++ww) s += nxt[ww][tid];
    if (s) atomicAdd(hist_next + tid, s);
  }
}

// ---- global histogram of one key byte: thread-private byte counters in a bank-conflict-free layout ------------------------
// The [digit][thread] byte table of radix_upsweep_kernel puts a warp's 32 increments on ~random banks (3.4 wavefronts per LDS.U8 and
// per STS.U8: the kernel runs at the shared-memory port, 0.17 ms per 100 M keys).  Here thread t owns WORD t of each of 64 rows; row g
// packs its counters of the digits 4 g .. 4 g + 3 as the four bytes of that word, so every access of a warp hits bank = lane: one
// wavefront per instruction.  A thread adds at most 252 keys between two flushes; a flush sums each row's bytes with dp4a (thread
// (g, q) takes the quarter row q, skewed by its lane so that the 32 lanes again read 32 different banks) and clears the table.
constexpr int kPrivBytes = 64 * 256 * 4;

__device__ __forceinline__ void priv_count(uint32_t* tab, uint32_t d) {
  uint32_t* w = tab + (d >> 2) * 256 + threadIdx.x;
  *w += 1u << ((d & 3u) * 8u);
}
// called by all 256 threads between two barriers; tot[k] of the threads with (tid & 3) == 0 belongs to digit 4 (tid >> 2) + k
__device__ __forceinline__ void priv_flush(uint32_t* tab, uint32_t (&tot)[4]) {
  const int lane = threadIdx.x & 31;
  uint32_t* row = tab + (threadIdx.x >> 2) * 256 + (threadIdx.x & 3) * 64;
  uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll 8
  for (int j = 0; j < 64; ++j) {
    const int jj = (j + lane) & 63;
    const uint32_t w = row[jj];
    row[jj] = 0u;
    s0 = __dp4a(w, 0x00000001u, s0);
    s1 = __dp4a(w, 0x00000100u, s1);
    s2 = __dp4a(w, 0x00010000u, s2);
    s3 = __dp4a(w, 0x01000000u, s3);
  }
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    s3 += __shfl_xor_sync(0xffffffffu, s3, o);
  }
  tot[0] += s0; tot[1] += s1; tot[2] += s2; tot[3] += s3;
}
__device__ __forceinline__ void priv_publish(const uint32_t (&tot)[4], uint32_t* __restrict__ hist) {
  if ((threadIdx.x & 3) == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (tot[k]) atomicAdd(hist + (threadIdx.x >> 2) * 4 + k, tot[k]);
  }
}

// used when the prep kernel's histogram of byte 0 is not the first pass's (byte 0 constant over all keys) and, in histogram mode 0,
// before every later pass
template <typename K>
__global__ void __launch_bounds__(256) radix_hist_kernel(const K* __restrict__ kin, long long n, int shift, uint32_t* __restrict__ hist) {
  extern __shared__ __align__(16) uint32_t tab[];   // [64][256]
  constexpr int V = VecOf<K>::V, U = 4;              // U 16-byte loads in flight per thread (the workspace keys are 256-byte aligned)
  for (int i = threadIdx.x; i < 64 * 256; i += 256) tab[i] = 0u;
  __syncthreads();
  const long long stride = (long long)gridDim.x * (256 * V * U);
  const long long iters = (n + stride - 1) / stride;
  uint32_t tot[4] = {0u, 0u, 0u, 0u};
  int since = 0;
  for (long long it = 0; it < iters; ++it) {
    const long long b = it * stride + (long long)blockIdx.x * (256 * V * U) + (long long)threadIdx.x * V;
    K k[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = b + (long long)u * 256 * V;
      if (i + V <= n) {
        load_vec<K, V>(kin + i, k[u]);
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) k[u][e] = i + e < n ? kin[i + e] : (K)0;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = b + (long long)u * 256 * V;
#pragma unroll
      for (int e = 0; e < V; ++e)
        if (i + e < n) priv_count(tab, (uint32_t)(k[u][e] >> shift) & 0xffu);
    }
    since += U * V;
    if (since + U * V > 255 || it == iters - 1) {   // block-uniform
      __syncthreads();
      priv_flush(tab, tot);
      __syncthreads();
      since = 0;
    }
  }
  priv_publish(tot, hist);
}

// sort_prep_kernel + the global histogram of key byte 0 in the same read of the scores; WRITE = false: header and histogram only (the
// first radix pass then reads the scores itself)
template <typename K, bool WRITE>
__global__ void __launch_bounds__(256) sort_prep_hist_kernel(const typename ScoreOf<K>::type* __restrict__ scores, const uint8_t* __restrict__ labels,
                                                              long long n, K* __restrict__ keys, uint32_t* __restrict__ pay, SortHeader* __restrict__ hdr,
                                                              uint32_t* __restrict__ hist0) {
  typedef typename ScoreOf<K>::type S;
  extern __shared__ __align__(16) uint32_t tab[];   // [64][256]
  for (int i = threadIdx.x; i < 64 * 256; i += 256) tab[i] = 0u;
  __syncthreads();
  constexpr int U = WRITE ? 8 : 16;   // scores and labels in flight per thread (scalar loads: the caller's pointers need no alignment)
  const long long stride = (long long)gridDim.x * (256 * U);
  const long long iters = (n + stride - 1) / stride;
  uint32_t ones = 0;
  uint32_t tot[4] = {0u, 0u, 0u, 0u};
  K kand = ~(K)0, kor = 0;
  int since = 0;
  for (long long it = 0; it < iters; ++it) {
    const long long b = it * stride + (long long)blockIdx.x * (256 * U) + threadIdx.x;
    S s[U];
    uint8_t lb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = b + u * 256;
      s[u] = i < n ? scores[i] : (S)0;
      lb[u] = i < n ? labels[i] : (uint8_t)0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = b + u * 256;
      if (i < n) {
        const K k = to_key(s[u]);
        const uint32_t lab = lb[u] != 0;
        if constexpr (WRITE) {
          keys[i] = k;
          pay[i] = (uint32_t)i | (lab << 31);
        }
        ones += lab;
        kand &= k;
        kor |= k;
        priv_count(tab, (uint32_t)k & 0xffu);
      }
    }
    since += U;
    if (since + U > 255 || it == iters - 1) {   // block-uniform
      __syncthreads();
      priv_flush(tab, tot);
      __syncthreads();
      since = 0;
    }
  }
  priv_publish(tot, hist0);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    ones += __shfl_xor_sync(0xffffffffu, ones, o);
    kand &= (K)__shfl_xor_sync(0xffffffffu, (unsigned long long)kand, o);
    kor |= (K)__shfl_xor_sync(0xffffffffu, (unsigned long long)kor, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (ones) atomicAdd(&hdr->ones, (unsigned long long)ones);
    atomicAnd(&hdr->key_and, (unsigned long long)kand);
    atomicOr(&hdr->key_or, (unsigned long long)kor);
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

// single-block exclusive scan of the per-tile bonafide counts (<= ~260k tiles at n = 2^30).  The counts of 16 chunks of 1,024 tiles are
// loaded together before the dependent chain of block scans starts (the first version paid one L2 round trip per chunk: 36 us at 100 M).
// cross_tile != nullptr (full sweep: curve points 0 .. n): the tile whose first curve point is >= 0 and whose last one is < 0 is recorded
// (the test of sweep_min_kernel, same fp64 operations; FAR - FRR does not increase along the curve), so that the arg-min kernel runs as ONE CTA instead of one early-exiting CTA per tile.
__global__ void __launch_bounds__(1024) sweep_scan_kernel(const uint32_t* __restrict__ block_ones, long long nb,
                                                           unsigned long long* __restrict__ block_excl, long long n, long long n_bona,
                                                           long long n_spoof, int* __restrict__ cross_tile) {
  __shared__ unsigned long long wsum[2][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (cross_tile != nullptr && threadIdx.x == 0) *cross_tile = 0x7fffffff;
  __syncthreads();
  unsigned long long carry = 0;   // kept identically by every thread
  int par = 0;
  for (long long start0 = 0; start0 < nb; start0 += 16 * 1024) {
    uint32_t v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const long long i = start0 + c * 1024 + threadIdx.x;
      v[c] = i < nb ? block_ones[i] : 0u;
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const long long start = start0 + c * 1024;
      if (start >= nb) break;   // block-uniform
      const long long i = start + threadIdx.x;
      unsigned long long incl = v[c];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      if (lane == 31) wsum[par][w] = incl;
      __syncthreads();
      const unsigned long long mine = wsum[par][lane];   // warp totals, one per lane
      unsigned long long winc = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += up;
      }
      const unsigned long long total = __shfl_sync(0xffffffffu, winc, 31);
      const unsigned long long woff = __shfl_sync(0xffffffffu, winc - mine, w);
      if (i < nb) {
        const unsigned long long ex = carry + woff + incl - v[c];
        block_excl[i] = ex;
        if (cross_tile != nullptr) {   // a tile's first point is the last point of the tile before it (+1 at k = 0): the first tile ending below 0
          const long long ks = i * kSortTile, len = (n - ks) < kSortTile ? (n - ks) : kSortTile;
          const long long ke = ks + len, c1e = (long long)ex + (long long)v[c];
          const double de = __dsub_rn(__ddiv_rn((double)(n_spoof - (ke - c1e)), (double)n_spoof), __ddiv_rn((double)c1e, (double)n_bona));
          if (de < 0.0) atomicMin(cross_tile, (int)i);
        }
      }
      carry += total;
      par ^= 1;   // the next chunk writes the other half of wsum: one barrier per chunk
    }
  }
}

__device__ __forceinline__ bool better(double d, long long i, double bd, long long bi) { return d < bd || (d == bd && i < bi); }

// The curve points evaluated are k = k_base .. k_base + n, where the first k_base sorted scores (c1_base of them bonafide)
// precede pay[0] (k_base = c1_base = 0 for the full sweep; the radix-select path sweeps one tie group only).
// FAR - FRR is strictly decreasing in k, so the arg-min of |FAR - FRR| is the last non-negative or the first negative point:
// only the ONE tile whose first point is >= 0 and whose last point is < 0 can hold it (both computed from the tile's label
// counts); every other tile returns at once without reading its payloads or doing fp64 divisions.
__global__ void __launch_bounds__(256) sweep_min_kernel(const uint32_t* __restrict__ pay, long long n, long long n_bona, long long n_spoof,
                                                         const unsigned long long* __restrict__ block_excl, const uint32_t* __restrict__ block_ones,
                                                         SweepBest* __restrict__ block_best, long long k_base, long long c1_base, int single,
                                                         const int* __restrict__ cross_tile) {
  // cross_tile (with single): the tile that holds the crossing, found by sweep_scan_kernel -- one CTA is launched instead of one per tile
  const long long bx = cross_tile ? (long long)*cross_tile : (long long)blockIdx.x;
  // single != 0: exactly one tile holds the crossing (the curve runs from +1 at k = 0 to -1 at k = n); it alone writes block_best[0]
  {
    const long long t0 = bx * kSortTile;
    const long long len = (n - t0) < kSortTile ? (n - t0) : kSortTile;
    const long long ks = k_base + t0, c1s = c1_base + (long long)block_excl[bx];
    const long long ke = ks + len, c1e = c1s + (long long)block_ones[bx];
    const double ds = __dsub_rn(__ddiv_rn((double)(n_spoof - (ks - c1s)), (double)n_spoof), __ddiv_rn((double)c1s, (double)n_bona));
    const double de = __dsub_rn(__ddiv_rn((double)(n_spoof - (ke - c1e)), (double)n_spoof), __ddiv_rn((double)c1e, (double)n_bona));
    if (!(ds >= 0.0 && de < 0.0)) {   // block-uniform
      if (threadIdx.x == 0 && !single) block_best[bx] = SweepBest{1.0e300, 0x7fffffffffffffffll, 0};
      return;
    }
  }
  // blocked arrangement: thread t owns sorted positions base + 16 t .. + 15
  const long long base = bx * kSortTile + (long long)threadIdx.x * kSortItems;
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
  long long c1 = c1_base + (long long)block_excl[bx] + woff + incl - ones;

  const double dspoof = (double)n_spoof, dbona = (double)n_bona;
  double bd = 1.0e300;
  long long bi = 0x7fffffffffffffffll, bc1 = 0;
  if (threadIdx.x == 0) {  // the point just before this tile's first score (k = 0 for the first tile: FAR = 1, FRR = 0)
    const long long ks = k_base + bx * kSortTile;
    const double far0 = __ddiv_rn((double)(n_spoof - (ks - c1)), dspoof);
    const double frr0 = __ddiv_rn((double)c1, dbona);
    bd = fabs(__dsub_rn(far0, frr0));
    bi = ks;
    bc1 = c1;
  }
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const long long j = base + i;
    if (j < n) {
      c1 += lab[i];
      const long long k = k_base + j + 1;
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
    block_best[single ? 0 : bx] = b;
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

// one radix pass of the one-sweep form: status words cleared, kernel launched on as many CTAs as are resident at once
template <typename K, int HIST, int T, bool FIRST>
static int onesweep_launch(const K* kin, const uint32_t* pin, K* kout, uint32_t* pout, int64_t n, int shift, int next_shift, const uint32_t* hist_cur,
                           uint32_t* hist_next, uint32_t* status, unsigned int* ticket, int num_sms, cudaStream_t stream, bool clear_status = true) {
  typedef OsCfg<K, T> Cfg;
  auto kern = radix_onesweep_kernel<K, HIST, T, FIRST>;
  static bool configured[32] = {false};
  if (dfs_first_use_on_device(configured)) DFS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  int ctas = 0;
  DFS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, T, Cfg::SMEM));
  DFS_REQUIRE(ctas >= 1, DFS_ERR_CUDA, "dfs_eer: the one-sweep kernel does not fit on this device");
  const long long tiles = ceil_div64(n, Cfg::TILE);
  if (clear_status) DFS_CUDA_CHECK(cudaMemsetAsync(status, 0, (size_t)tiles * 256 * 4, stream));
  kern<<<(unsigned)std::min<long long>(tiles, (long long)num_sms * ctas), T, Cfg::SMEM, stream>>>(
      kin, pin, kout, pout, (uint32_t)n, shift, next_shift, hist_cur, hist_next, status, ticket);
  DFS_LAUNCH_CHECK();
  return DFS_OK;
}
// dfs_set_global_option("eer_sort_onesweep"): form of dfs_eer's radix passes
//   0 = count / scan / scatter kernels over per-CTA super-tiles (round 1; cross-check)
//   1 = one-sweep, 512-thread tiles, next pass's histogram by shared-memory atomics in the scatter kernel (default)
//   2 = as 1 on 256-thread tiles      3 = as 1, histogram by a kernel of its own before each pass      4 = as 1, histogram by ballots
//   5 = as 1, the first pass reads the scores and labels itself (no key / payload image written before it)
int g_sort_onesweep = 1;
// dfs_set_global_option("eer_sort_overlap") (fp32 scores, one-sweep forms other than 3 and 5): 1 = the pass over key byte 0 is launched
// BEFORE the host reads the header back (label count, key AND / OR), on the assumption that byte 0 varies and byte 1 is sorted next; the
// header copy runs on a side stream behind the prep kernel, so the GPU does not idle for the host round trip, and the status words of
// all passes are cleared by one memset ahead of the prep kernel instead of one between every two passes (2.686 -> 2.669 ms per 100 M).
// If byte 0 turns out constant, that pass was an identity permutation (one wasted pass on such inputs, same result); if byte 1 is
// constant, the next pass's histogram is taken by radix_hist_kernel.  0 = header first, then the passes (cross-check).
int g_sort_overlap = 1;
struct SortSideStream {
  cudaStream_t copy = nullptr;
  cudaEvent_t prepped = nullptr, copied = nullptr;
  unsigned long long* pinned = nullptr;   // SortHeader lands here
};
static SortSideStream g_side[32];
template <typename K, typename... A>
static int onesweep_pass(int form, A... a) {
  switch (form) {
    case 2: return onesweep_launch<K, 1, 256, false>(a...);
    case 3: return onesweep_launch<K, 0, 512, false>(a...);
    case 4: return onesweep_launch<K, 2, 512, false>(a...);
    default: return onesweep_launch<K, 1, 512, false>(a...);
  }
}

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
  const long long supers = ceil_div64(n, kSuperKeys);
  const size_t o_counts = carve((size_t)supers * 256 * 4);
  const size_t o_hist = carve(3 * 256 * 4);   // bucket bases of the current pass | per-digit totals | completion counter of the scan
  // one-sweep form: global histogram of every key byte [8][256] + one tile ticket per pass [8] (zeroed together), look-back status words
  const long long os_tiles = ceil_div64(n, OsCfg<K, 256>::TILE);   // the smaller of the two tile sizes
  const size_t o_oshist = carve((8 * 256 + 8) * 4);
  const size_t status_stride = align_up((size_t)os_tiles * 256 * 4, 256);   // fp32 keys: one region per pass (cleared together in the overlap form)
  const size_t o_status = carve(status_stride * (sizeof(K) == 4 ? PASSES : 1));
  const size_t o_small = carve(256);      // SortHeader
  const size_t o_bones = carve((size_t)tiles * 4), o_bexcl = carve((size_t)tiles * 8), o_bbest = carve((size_t)tiles * sizeof(SweepBest));
  const size_t o_res = carve(sizeof(dfs_eer_result));
  void* base = nullptr;
  DFS_PROPAGATE(get_workspace(off, &base));
  uint8_t* b8 = static_cast<uint8_t*>(base);
  K* keys[2] = {reinterpret_cast<K*>(b8 + o_k0), reinterpret_cast<K*>(b8 + o_k1)};
  uint32_t* pay[2] = {reinterpret_cast<uint32_t*>(b8 + o_p0), reinterpret_cast<uint32_t*>(b8 + o_p1)};
  uint32_t* counts = reinterpret_cast<uint32_t*>(b8 + o_counts);
  uint32_t* hist = reinterpret_cast<uint32_t*>(b8 + o_hist);
  SortHeader* hdr = reinterpret_cast<SortHeader*>(b8 + o_small);
  uint32_t* bones = reinterpret_cast<uint32_t*>(b8 + o_bones);
  unsigned long long* bexcl = reinterpret_cast<unsigned long long*>(b8 + o_bexcl);
  SweepBest* bbest = reinterpret_cast<SweepBest*>(b8 + o_bbest);
  dfs_eer_result* res_dev = reinterpret_cast<dfs_eer_result*>(b8 + o_res);

  uint32_t* oshist = reinterpret_cast<uint32_t*>(b8 + o_oshist);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(oshist + 8 * 256);
  uint32_t* status = reinterpret_cast<uint32_t*>(b8 + o_status);
  const bool onesweep = g_sort_onesweep != 0;

  // SortHeader{ones = 0, key_and = ~0, key_or = 0} by memsets: no host buffer to keep alive, no synchronisation before the first kernel
  DFS_CUDA_CHECK(cudaMemsetAsync(hdr, 0, sizeof(SortHeader), stream));
  DFS_CUDA_CHECK(cudaMemsetAsync(&hdr->key_and, 0xff, sizeof(unsigned long long), stream));
  DFS_CUDA_CHECK(cudaMemsetAsync(hist + 512, 0, 4, stream));   // completion counter of radix_scan_kernel (the workspace is shared and grow-only)
  int num_sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  static bool configured[32] = {false};
  const size_t dyn_smem = (sizeof(K) + 4) * kSortTile;
  if (dfs_first_use_on_device(configured)) {
    DFS_CUDA_CHECK(cudaFuncSetAttribute(radix_downsweep_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(radix_upsweep_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * kCntStride));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(radix_hist_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPrivBytes));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(sort_prep_hist_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPrivBytes));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(sort_prep_hist_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPrivBytes));
  }
  const bool fused_first = g_sort_onesweep == 5;
  const unsigned hist_grid = (unsigned)std::min<long long>(ceil_div64(n, 2048), (long long)num_sms * 3);   // 3 CTAs of 64 KB per SM
  // overlap form (see g_sort_overlap): pass 0 speculatively ahead of the header read-back
  const bool overlap = onesweep && g_sort_overlap != 0 && sizeof(K) == 4 && g_sort_onesweep != 3 && !fused_first;
  auto status_of = [&](int ps) { return reinterpret_cast<uint32_t*>(b8 + o_status + (overlap ? (size_t)ps * status_stride : 0)); };
  SortSideStream* side = nullptr;
  if (overlap) {
    int dev = 0;
    cudaGetDevice(&dev);
    DFS_REQUIRE(dev >= 0 && dev < 32, DFS_ERR_CUDA, "dfs_eer: device index out of range");
    side = &g_side[dev];
    if (side->copy == nullptr) {
      DFS_CUDA_CHECK(cudaStreamCreateWithFlags(&side->copy, cudaStreamNonBlocking));
      DFS_CUDA_CHECK(cudaEventCreateWithFlags(&side->prepped, cudaEventDisableTiming));
      DFS_CUDA_CHECK(cudaEventCreateWithFlags(&side->copied, cudaEventDisableTiming));
      DFS_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&side->pinned), sizeof(SortHeader), cudaHostAllocDefault));
    }
    DFS_CUDA_CHECK(cudaMemsetAsync(b8 + o_status, 0, status_stride * PASSES, stream));
  }
  if (onesweep) {
    DFS_CUDA_CHECK(cudaMemsetAsync(oshist, 0, (8 * 256 + 8) * 4, stream));
    if (fused_first)
      sort_prep_hist_kernel<K, false><<<hist_grid, 256, kPrivBytes, stream>>>(static_cast<const S*>(scores), labels, n, nullptr, nullptr, hdr, oshist);
    else
      sort_prep_hist_kernel<K, true><<<hist_grid, 256, kPrivBytes, stream>>>(static_cast<const S*>(scores), labels, n, keys[0], pay[0], hdr, oshist);
  } else {
    const unsigned prep_grid = (unsigned)std::min<long long>(ceil_div64(n, 256), (long long)num_sms * 8);
    sort_prep_kernel<K><<<prep_grid, 256, 0, stream>>>(static_cast<const S*>(scores), labels, n, keys[0], pay[0], hdr);
  }
  DFS_LAUNCH_CHECK();
  // the label count and the key AND/OR decide the host control flow (single-class early-out; skipped passes)
  SortHeader small_host;
  int cur = 0;
  if (overlap) {
    DFS_CUDA_CHECK(cudaEventRecord(side->prepped, stream));
    DFS_CUDA_CHECK(cudaStreamWaitEvent(side->copy, side->prepped, 0));
    DFS_CUDA_CHECK(cudaMemcpyAsync(side->pinned, hdr, sizeof(SortHeader), cudaMemcpyDeviceToHost, side->copy));
    DFS_CUDA_CHECK(cudaEventRecord(side->copied, side->copy));
    // byte 0 -> buffer 1, histogram of byte 1 on the way (PASSES == 4 here)
    DFS_PROPAGATE(onesweep_pass<K>(g_sort_onesweep, (const K*)keys[0], (const uint32_t*)pay[0], keys[1], pay[1], n, 0, 8, (const uint32_t*)oshist, oshist + 256,
                                   status_of(0), tickets + 0, num_sms, stream, false));
    cur = 1;
    DFS_CUDA_CHECK(cudaEventSynchronize(side->copied));
    memcpy(&small_host, side->pinned, sizeof(small_host));
  } else {
    DFS_CUDA_CHECK(cudaMemcpyAsync(&small_host, hdr, sizeof(small_host), cudaMemcpyDeviceToHost, stream));
    DFS_CUDA_CHECK(cudaStreamSynchronize(stream));
  }
  const long long n_bona = (long long)small_host.ones, n_spoof = n - n_bona;

  if (onesweep) {
    int pass_list[8], np = 0;
    for (int ps = overlap ? 1 : 0; ps < PASSES; ++ps)   // overlap: byte 0 is sorted already (an identity pass if it was constant)
      if ((((small_host.key_and ^ small_host.key_or) >> (8 * ps)) & 0xffull) != 0) pass_list[np++] = ps;   // other bytes: identity passes
    const bool hist_kernel = g_sort_onesweep == 3;
    // the first pass can read the scores itself when it sorts byte 0 (whose histogram the header kernel took); otherwise (byte 0 constant
    // over all scores, or every score equal) the key / payload image is written after all and the passes start from it
    bool image = !fused_first;
    if (fused_first && (np == 0 || pass_list[0] != 0)) {
      const unsigned prep_grid = (unsigned)std::min<long long>(ceil_div64(n, 256), (long long)num_sms * 8);
      sort_prep_kernel<K><<<prep_grid, 256, 0, stream>>>(static_cast<const S*>(scores), labels, n, keys[0], pay[0], hdr);   // hdr: counted twice, not read again
      DFS_LAUNCH_CHECK();
      image = true;
    }
    // whose histogram exists already: byte 0 (prep kernel), or byte 1 (the speculative pass) in the overlap form
    if (np > 0 && pass_list[0] != (overlap ? 1 : 0)) {
      radix_hist_kernel<K><<<hist_grid, 256, kPrivBytes, stream>>>(keys[cur], n, 8 * pass_list[0], oshist + 256 * pass_list[0]);
      DFS_LAUNCH_CHECK();
    }
    for (int ip = 0; ip < np; ++ip) {
      const int ps = pass_list[ip], nx = ip + 1 < np ? pass_list[ip + 1] : -1;
      if (hist_kernel && ip > 0) {
        radix_hist_kernel<K><<<hist_grid, 256, kPrivBytes, stream>>>(keys[cur], n, 8 * ps, oshist + 256 * ps);
        DFS_LAUNCH_CHECK();
      }
      if (ip == 0 && !image) {   // scores -> buffer 0
        DFS_PROPAGATE((onesweep_launch<K, 1, 512, true>(static_cast<const K*>(scores), reinterpret_cast<const uint32_t*>(labels), keys[0], pay[0], n, 8 * ps,
                                                        nx < 0 ? -1 : 8 * nx, oshist + 256 * ps, oshist + 256 * (nx < 0 ? 0 : nx), status, tickets + ps, num_sms,
                                                        stream)));
        cur = 0;
        continue;
      }
      DFS_PROPAGATE(onesweep_pass<K>(g_sort_onesweep, (const K*)keys[cur], (const uint32_t*)pay[cur], keys[cur ^ 1], pay[cur ^ 1], n, 8 * ps, nx < 0 ? -1 : 8 * nx,
                                     (const uint32_t*)(oshist + 256 * ps), oshist + 256 * (nx < 0 ? 0 : nx), status_of(ps), tickets + ps, num_sms, stream, !overlap));
      cur ^= 1;
    }
  }
  for (int ps = 0; ps < PASSES && !onesweep; ++ps) {
    if ((((small_host.key_and ^ small_host.key_or) >> (8 * ps)) & 0xffull) == 0) continue;  // every key shares this digit: identity pass
    radix_upsweep_kernel<K><<<(unsigned)supers, 256, 256 * kCntStride, stream>>>(keys[cur], n, 8 * ps, counts);
    DFS_LAUNCH_CHECK();
    radix_scan_kernel<<<256, 256, 0, stream>>>(counts, (int)supers, hist, hist + 256, reinterpret_cast<unsigned int*>(hist + 512));
    DFS_LAUNCH_CHECK();
    radix_downsweep_kernel<K><<<(unsigned)supers, kSortThreads, dyn_smem, stream>>>(keys[cur], pay[cur], keys[cur ^ 1], pay[cur ^ 1], n, 8 * ps,
                                                                                    hist, counts);
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
  int* cross = reinterpret_cast<int*>(tickets);   // the pass tickets are spent
  sweep_scan_kernel<<<1, 1024, 0, stream>>>(bones, tiles, bexcl, n, n_bona, n_spoof, cross);
  DFS_LAUNCH_CHECK();
  sweep_min_kernel<<<1, 256, 0, stream>>>(pay[cur], n, n_bona, n_spoof, bexcl, bones, bbest, 0, 0, /*single=*/1, cross);
  DFS_LAUNCH_CHECK();
  sweep_final_kernel<K><<<1, 256, 0, stream>>>(bbest, 1, keys[cur], n, n_bona, n_spoof, res_dev);
  DFS_LAUNCH_CHECK();
  DFS_CUDA_CHECK(cudaMemcpyAsync(result_host, res_dev, sizeof(dfs_eer_result), cudaMemcpyDeviceToHost, stream));
  DFS_CUDA_CHECK(cudaStreamSynchronize(stream));
  return DFS_OK;
}

// ---- EER by radix SELECT (no permutation requested) ------------------------------------------------------
// FAR(k) - FRR(k) is strictly decreasing in k (every sorted score either removes 1/n_spoof from FAR or adds 1/n_bona to
// FRR, both far above an fp64 ulp), so argmin |FAR - FRR| is the last k with a non-negative difference or the first
// with a negative one.  That crossing is found like a quantile: an MSD radix descent over the key bytes with one
// (digit, label) histogram per level -- 5 B/score per level instead of a full sort -- which pins down the key value G
// holding the crossing, the label counts below G and the size of G's tie group.  A single-label tie group (always the
// case for tie-free scores) is resolved in closed form; a mixed-label group is compacted in original index order (the
// stable-sort contract) and swept with the same kernels as the full sort.  Every FAR/FRR value is produced by the same
// IEEE fp64 operations as in sweep_min_kernel, so (eer, threshold, eer_idx) are bit-identical to the sort path.
struct SelectState {
  unsigned long long prefix;     // key bits decided so far
  unsigned long long mask;       // which bits those are
  unsigned long long c0_below;   // spoof / bonafide scores with key below the current prefix range
  unsigned long long c1_below;
  unsigned long long n_bona, n_spoof;
  unsigned long long key_and, key_or;
  unsigned long long g0, g1;     // spoof / bonafide scores inside the current prefix range
  unsigned long long pred;       // largest key below G (filled on demand)
  SweepBest best;                // chosen curve point
  int status;                    // 1 = single-class input
  int resolved;                  // best is valid
};

__global__ void select_init_kernel(SelectState* st, unsigned long long* hist) {
  if (threadIdx.x == 0) {
    SelectState z{};
    z.key_and = ~0ull;
    *st = z;
  }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) hist[i] = 0ull;
}

constexpr int kSelIters = 56;   // vector loads per thread between two flushes: <= 224 keys per thread
constexpr int kSelUnroll = 8;

template <typename K, bool FIRST>
__global__ void __launch_bounds__(256, 1) select_hist_kernel(const typename ScoreOf<K>::type* __restrict__ scores,
                                                              const uint8_t* __restrict__ labels, long long n, int shift,
                                                              SelectState* __restrict__ st, unsigned long long* __restrict__ hist /*[2][256]*/) {
  typedef typename ScoreOf<K>::type S;
  constexpr int V = VecOf<K>::V;
  extern __shared__ __align__(16) uint8_t cnt[];   // [512 = label*256 + digit][kCntStride]
  for (int i = threadIdx.x; i < 512 * kCntStride / 4; i += 256) reinterpret_cast<uint32_t*>(cnt)[i] = 0u;
  const K prefix = FIRST ? (K)0 : (K)st->prefix;
  const K mask = FIRST ? (K)0 : (K)st->mask;
  __syncthreads();
  uint8_t* mine = cnt + threadIdx.x;
  K kand = ~(K)0, kor = 0;
  unsigned long long tot0 = 0, tot1 = 0;
  constexpr long long kChunk = 256ll * V * kSelIters;
  const long long nvec = n - (n % V);   // the last n % V scores are handled by one thread below
  for (long long c0 = (long long)blockIdx.x * kChunk; c0 < nvec; c0 += (long long)gridDim.x * kChunk) {
    for (int it0 = 0; it0 < kSelIters; it0 += kSelUnroll) {
      S sv[kSelUnroll][V];
      uint32_t lv[kSelUnroll];
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u) {
        const long long i = c0 + ((long long)(it0 + u) * 256 + threadIdx.x) * V;
        if (i < nvec) {
          load_vec<S, V>(scores + i, sv[u]);
          if constexpr (V == 4) lv[u] = __ldg(reinterpret_cast<const uint32_t*>(labels + i));
          else lv[u] = __ldg(reinterpret_cast<const uint16_t*>(labels + i));
        } else {
          lv[u] = 0u;
#pragma unroll
          for (int e = 0; e < V; ++e) sv[u][e] = (S)0;
        }
      }
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u) {
        const long long i = c0 + ((long long)(it0 + u) * 256 + threadIdx.x) * V;
        if (i < nvec) {
#pragma unroll
          for (int e = 0; e < V; ++e) {
            const K k = to_key(sv[u][e]);
            const uint32_t lab = ((lv[u] >> (8 * e)) & 0xffu) != 0u;
            if (FIRST) { kand &= k; kor |= k; }
            if (FIRST || (k & mask) == prefix) mine[(((uint32_t)(k >> shift) & 0xffu) | (lab << 8)) * kCntStride] += 1;
          }
        }
      }
    }
    __syncthreads();
    tot0 += flush_counter_row(cnt, threadIdx.x);
    tot1 += flush_counter_row(cnt, 256 + threadIdx.x);
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long i = nvec; i < n; ++i) {
      const K k = to_key(scores[i]);
      const uint32_t lab = labels[i] != 0;
      if (FIRST) { kand &= k; kor |= k; }
      if (FIRST || (k & mask) == prefix) atomicAdd(&hist[(((uint32_t)(k >> shift) & 0xffu)) + 256 * lab], 1ull);
    }
  }
  if (tot0) atomicAdd(&hist[threadIdx.x], tot0);
  if (tot1) atomicAdd(&hist[256 + threadIdx.x], tot1);
  if (FIRST) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      kand &= (K)__shfl_xor_sync(0xffffffffu, (unsigned long long)kand, o);
      kor |= (K)__shfl_xor_sync(0xffffffffu, (unsigned long long)kor, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAnd(&st->key_and, (unsigned long long)kand);   // zero-extended: the upper bits of a 32-bit key AND to 0, OR to 0
      atomicOr(&st->key_or, (unsigned long long)kor);
    }
  }
}

// TMA-fed variant (16-byte aligned score / label pointers): one producer lane keeps a ring of kSelStages stages, each one
// 16 KB of scores + their labels, in flight with cp.async.bulk (~80 KB per SM: enough outstanding bytes to cover HBM
// latency with a single resident CTA -- the counter table takes 130 KB of the SM's shared memory).  256 consumer threads
// read the stage with conflict-free 16-byte LDS and bump their private counters.
constexpr int kSelStages = 4;
constexpr int kSelStageB = 16384 + 4096;
constexpr int kSelRingOff = 513 * kCntStride + 60;                       // 133,440: 513 rows (row 512 = spare), 128-byte aligned
constexpr int kSelBarOff = kSelRingOff + kSelStages * kSelStageB;        // 215,360
constexpr int kSelTmaSmem = kSelBarOff + 2 * kSelStages * 8;

template <typename K, bool FIRST>
__global__ void __launch_bounds__(288, 1) select_hist_tma_kernel(const typename ScoreOf<K>::type* __restrict__ scores,
                                                                  const uint8_t* __restrict__ labels, long long n, int shift,
                                                                  SelectState* __restrict__ st, unsigned long long* __restrict__ hist /*[2][256]*/) {
  typedef typename ScoreOf<K>::type S;
  constexpr int V = VecOf<K>::V;
  constexpr int KS = 16384 / (int)sizeof(S);          // scores per stage: 4096 (fp32) / 2048 (fp64)
  constexpr int ITERS = KS / (256 * V);               // 16-byte loads per consumer thread per stage (4)
  constexpr int FLUSH_STAGES = 224 / (ITERS * V);     // <= 224 counts per thread between flushes
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* cnt = smem;
  uint8_t* ring = smem + kSelRingOff;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSelBarOff);
  uint64_t* empty = full + kSelStages;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 513 * kCntStride / 4; i += 288) reinterpret_cast<uint32_t*>(cnt)[i] = 0u;
  if (tid == 0) {
    for (int i = 0; i < kSelStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
    fence_mbar_init();
  }
  const K prefix = FIRST ? (K)0 : (K)st->prefix;
  const K mask = FIRST ? (K)0 : (K)st->mask;
  __syncthreads();
  const long long n_stages = n / KS;                  // full stages; the tail (< KS scores) is read directly below
  if (warp == 8) {
    // ===================== producer =====================
    if ((tid & 31) == 0) {
      uint32_t it = 0;
      for (long long g = blockIdx.x; g < n_stages; g += gridDim.x, ++it) {
        const int s = it % kSelStages;
        mbar_wait(&empty[s], ((it / kSelStages) & 1) ^ 1, 31);
        mbar_arrive_expect_tx(&full[s], 16384 + KS);
        bulk_g2s(ring + s * kSelStageB, scores + g * KS, 16384, &full[s]);
        bulk_g2s(ring + s * kSelStageB + 16384, labels + g * KS, KS, &full[s]);
      }
    }
    return;
  }
  // ===================== consumers (warps 0..7) =====================
  uint8_t* mine = cnt + tid;
  K kand = ~(K)0, kor = 0;
  unsigned long long tot0 = 0, tot1 = 0;
  uint32_t it = 0;
  int since_flush = 0;
  auto flush = [&]() {
    asm volatile("bar.sync 1, 256;" ::: "memory");
    tot0 += flush_counter_row(cnt, tid);
    tot1 += flush_counter_row(cnt, 256 + tid);
    asm volatile("bar.sync 1, 256;" ::: "memory");
  };
  for (long long g = blockIdx.x; g < n_stages; g += gridDim.x, ++it) {
    const int s = it % kSelStages;
    mbar_wait(&full[s], (it / kSelStages) & 1, 32);
    const uint8_t* base = ring + s * kSelStageB;
#pragma unroll
    for (int i = 0; i < ITERS; ++i) {
      const int v = i * 256 + tid;                    // 16-byte vector index within the stage
      S sv[V];
      const uint4 q = *reinterpret_cast<const uint4*>(base + v * 16);
      if constexpr (V == 4) {
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) sv[e] = *reinterpret_cast<const S*>(&w[e]);
      } else {
        const uint64_t w[2] = {((uint64_t)q.y << 32) | q.x, ((uint64_t)q.w << 32) | q.z};
#pragma unroll
        for (int e = 0; e < 2; ++e) sv[e] = *reinterpret_cast<const S*>(&w[e]);
      }
      uint32_t lv;
      if constexpr (V == 4) lv = *reinterpret_cast<const uint32_t*>(base + 16384 + v * 4);
      else lv = *reinterpret_cast<const uint16_t*>(base + 16384 + v * 2);
      // Below the first level almost no score carries the chosen prefix: test that first (4-5 instructions per score)
      // and leave the counters alone unless some lane of the warp has a hit.  The update itself is branch-free per
      // score: one outside the prefix bumps the spare row 512.
      K kk[V];
      bool hit[V];
      bool any = FIRST;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        kk[e] = to_key(sv[e]);
        hit[e] = FIRST || (kk[e] & mask) == prefix;
        if (FIRST) { kand &= kk[e]; kor |= kk[e]; }
        any |= hit[e];
      }
      if (FIRST || __any_sync(0xffffffffu, any)) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const uint32_t lab = min((lv >> (8 * e)) & 0xffu, 1u);
          const uint32_t row = hit[e] ? (((uint32_t)(kk[e] >> shift) & 0xffu) | (lab << 8)) : 512u;
          mine[row * kCntStride] += 1;
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);      // this warp is done reading the stage
    if (++since_flush == FLUSH_STAGES) {
      flush();
      since_flush = 0;
    }
  }
  if (blockIdx.x == 0) {                               // tail: fewer than KS scores, strided direct loads (<= 16 per thread)
    for (long long i = n_stages * KS + tid; i < n; i += 256) {
      const K k = to_key(scores[i]);
      const uint32_t lab = labels[i] != 0;
      if (FIRST) { kand &= k; kor |= k; }
      if (FIRST || (k & mask) == prefix) atomicAdd(&hist[((uint32_t)(k >> shift) & 0xffu) + 256 * lab], 1ull);
    }
  }
  flush();
  if (tot0) atomicAdd(&hist[tid], tot0);
  if (tot1) atomicAdd(&hist[256 + tid], tot1);
  if (FIRST) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      kand &= (K)__shfl_xor_sync(0xffffffffu, (unsigned long long)kand, o);
      kor |= (K)__shfl_xor_sync(0xffffffffu, (unsigned long long)kor, o);
    }
    if ((tid & 31) == 0) {
      atomicAnd(&st->key_and, (unsigned long long)kand);
      atomicOr(&st->key_or, (unsigned long long)kor);
    }
  }
}

// scalar-load variant for misaligned score / label pointers (same counters, same flush rule)
template <typename K, bool FIRST>
__global__ void __launch_bounds__(256, 1) select_hist_scalar_kernel(const typename ScoreOf<K>::type* __restrict__ scores,
                                                                     const uint8_t* __restrict__ labels, long long n, int shift,
                                                                     SelectState* __restrict__ st, unsigned long long* __restrict__ hist) {
  extern __shared__ __align__(16) uint8_t cnt[];
  for (int i = threadIdx.x; i < 512 * kCntStride / 4; i += 256) reinterpret_cast<uint32_t*>(cnt)[i] = 0u;
  const K prefix = FIRST ? (K)0 : (K)st->prefix;
  const K mask = FIRST ? (K)0 : (K)st->mask;
  __syncthreads();
  uint8_t* mine = cnt + threadIdx.x;
  K kand = ~(K)0, kor = 0;
  unsigned long long tot0 = 0, tot1 = 0;
  constexpr long long kChunk = 256ll * 224;
  for (long long c0 = (long long)blockIdx.x * kChunk; c0 < n; c0 += (long long)gridDim.x * kChunk) {
    for (int it = 0; it < 224; ++it) {
      const long long i = c0 + (long long)it * 256 + threadIdx.x;
      if (i < n) {
        const K k = to_key(scores[i]);
        const uint32_t lab = labels[i] != 0;
        if (FIRST) { kand &= k; kor |= k; }
        if (FIRST || (k & mask) == prefix) mine[(((uint32_t)(k >> shift) & 0xffu) | (lab << 8)) * kCntStride] += 1;
      }
    }
    __syncthreads();
    tot0 += flush_counter_row(cnt, threadIdx.x);
    tot1 += flush_counter_row(cnt, 256 + threadIdx.x);
    __syncthreads();
  }
  if (tot0) atomicAdd(&hist[threadIdx.x], tot0);
  if (tot1) atomicAdd(&hist[256 + threadIdx.x], tot1);
  if (FIRST) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      kand &= (K)__shfl_xor_sync(0xffffffffu, (unsigned long long)kand, o);
      kor |= (K)__shfl_xor_sync(0xffffffffu, (unsigned long long)kor, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAnd(&st->key_and, (unsigned long long)kand);
      atomicOr(&st->key_or, (unsigned long long)kor);
    }
  }
}

__device__ __forceinline__ double eer_curve_diff(long long c0, long long c1, long long n_spoof, long long n_bona) {
  const double far = __ddiv_rn((double)(n_spoof - c0), (double)n_spoof);   // evaluation.py:21-23
  const double frr = __ddiv_rn((double)c1, (double)n_bona);                // evaluation.py:24-26
  return __dsub_rn(far, frr);
}

// One level of the descent: pick the first digit whose END lies beyond the crossing (difference < 0 after all its keys).
// skip != 0: every key shares this byte (key AND == OR); no histogram was taken.
__global__ void __launch_bounds__(256) select_pick_kernel(unsigned long long* __restrict__ hist, SelectState* __restrict__ st, int shift,
                                                           int first, int skip) {
  if (skip) {
    if (threadIdx.x == 0) {
      st->prefix |= st->key_and & (0xffull << shift);
      st->mask |= 0xffull << shift;
    }
    return;
  }
  __shared__ unsigned long long w0[8], w1[8];
  __shared__ int wfirst[8];
  __shared__ int single;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned long long c0 = hist[tid], c1 = hist[256 + tid];
  hist[tid] = 0ull;
  hist[256 + tid] = 0ull;
  unsigned long long i0 = c0, i1 = c1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long u0 = __shfl_up_sync(0xffffffffu, i0, o), u1 = __shfl_up_sync(0xffffffffu, i1, o);
    if (lane >= o) { i0 += u0; i1 += u1; }
  }
  if (lane == 31) { w0[w] = i0; w1[w] = i1; }
  __syncthreads();
  unsigned long long t0 = 0, t1 = 0;
  for (int ww = 0; ww < 8; ++ww) {
    if (ww < w) { i0 += w0[ww]; i1 += w1[ww]; }
    t0 += w0[ww];
    t1 += w1[ww];
  }
  if (first) {
    if (tid == 0) {
      st->n_spoof = t0;
      st->n_bona = t1;
      st->g0 = t0;
      st->g1 = t1;
      single = (t0 == 0 || t1 == 0);
      if (single) st->status = 1;
    }
    __syncthreads();
    if (single) return;
  }
  const long long n_spoof = (long long)(first ? t0 : st->n_spoof), n_bona = (long long)(first ? t1 : st->n_bona);
  if (!first && st->status != 0) return;
  const long long C0 = (long long)(st->c0_below + i0), C1 = (long long)(st->c1_below + i1);
  const bool neg = eer_curve_diff(C0, C1, n_spoof, n_bona) < 0.0;
  const uint32_t bal = __ballot_sync(0xffffffffu, neg);
  if (lane == 0) wfirst[w] = bal ? (32 * w + __ffs(bal) - 1) : 256;
  __syncthreads();
  int b = 256;
  for (int ww = 0; ww < 8; ++ww) b = wfirst[ww] < b ? wfirst[ww] : b;
  __syncthreads();   // everybody has read c*_below before the owner updates them
  if (tid == b) {
    st->c0_below += i0 - c0;
    st->c1_below += i1 - c1;
    st->prefix |= (unsigned long long)b << shift;
    st->mask |= 0xffull << shift;
    st->g0 = c0;
    st->g1 = c1;
  }
}

// After the last level: prefix = G.  A single-label tie group is resolved here (binary search for the first negative difference).
__global__ void select_uniform_group_kernel(SelectState* __restrict__ st) {
  if (threadIdx.x != 0 || st->status != 0) return;
  const long long g0 = (long long)st->g0, g1 = (long long)st->g1;
  if (g0 > 0 && g1 > 0) return;   // mixed labels: needs the compaction path
  const long long n_spoof = (long long)st->n_spoof, n_bona = (long long)st->n_bona;
  const long long b0 = (long long)st->c0_below, b1 = (long long)st->c1_below;
  const long long m = g0 + g1;
  const int lab = g1 > 0;
  // first j in [1, m] with diff(start + j) < 0 (exists: the pick step chose this group because its end is negative)
  long long lo = 1, hi = m;
  while (lo < hi) {
    const long long mid = lo + (hi - lo) / 2;
    const bool neg = eer_curve_diff(b0 + (lab ? 0 : mid), b1 + (lab ? mid : 0), n_spoof, n_bona) < 0.0;
    if (neg) hi = mid; else lo = mid + 1;
  }
  const long long j = lo;
  const long long c0a = b0 + (lab ? 0 : j - 1), c1a = b1 + (lab ? j - 1 : 0);
  const long long c0b = b0 + (lab ? 0 : j), c1b = b1 + (lab ? j : 0);
  const double da = fabs(eer_curve_diff(c0a, c1a, n_spoof, n_bona)), db = fabs(eer_curve_diff(c0b, c1b, n_spoof, n_bona));
  SweepBest best;
  if (da <= db) { best.diff = da; best.idx = c0a + c1a; best.c1 = c1a; }   // np.argmin keeps the lowest index among equals
  else { best.diff = db; best.idx = c0b + c1b; best.c1 = c1b; }
  st->best = best;
  st->resolved = 1;
}

// ---- mixed-label tie group: labels of the scores equal to G, in original index order ----
template <typename K>
__global__ void __launch_bounds__(256) group_count_kernel(const typename ScoreOf<K>::type* __restrict__ scores, long long n,
                                                           const SelectState* __restrict__ st, uint32_t* __restrict__ tile_cnt) {
  const K G = (K)st->prefix;
  const long long base = (long long)blockIdx.x * kSortTile;
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const long long j = base + threadIdx.x + i * 256;
    if (j < n) c += to_key(scores[j]) == G;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __shared__ uint32_t part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s += part[i];
    tile_cnt[blockIdx.x] = s;
  }
}

template <typename K>
__global__ void __launch_bounds__(256) group_write_kernel(const typename ScoreOf<K>::type* __restrict__ scores, const uint8_t* __restrict__ labels,
                                                           long long n, const SelectState* __restrict__ st,
                                                           const unsigned long long* __restrict__ tile_excl, uint32_t* __restrict__ pay_out) {
  const K G = (K)st->prefix;
  // blocked arrangement: thread t owns positions base + 16 t .. + 15, so ranks follow the original index order
  const long long base = (long long)blockIdx.x * kSortTile + (long long)threadIdx.x * kSortItems;
  uint32_t hit = 0, labs = 0;
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const long long j = base + i;
    if (j < n && to_key(scores[j]) == G) {
      hit |= 1u << i;
      labs |= (uint32_t)(labels[j] != 0) << i;
    }
  }
  const uint32_t mine = __popc(hit);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t incl = mine;
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
  unsigned long long dst = tile_excl[blockIdx.x] + woff + incl - mine;
#pragma unroll
  for (int i = 0; i < kSortItems; ++i)
    if (hit & (1u << i)) pay_out[dst++] = ((uint32_t)(base + i) & 0x7fffffffu) | (((labs >> i) & 1u) << 31);
}

__global__ void __launch_bounds__(256) select_reduce_best_kernel(const SweepBest* __restrict__ block_best, long long nb, SelectState* __restrict__ st) {
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
    st->best = b;
    st->resolved = 1;
  }
}

// largest key strictly below G: the threshold when the chosen curve point is the start of G's group (evaluation.py:37)
template <typename K>
__global__ void __launch_bounds__(256) select_pred_kernel(const typename ScoreOf<K>::type* __restrict__ scores, long long n, SelectState* __restrict__ st) {
  const K G = (K)st->prefix;
  K best = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const K k = to_key(scores[i]);
    if (k < G && k > best) best = k;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const K other = (K)__shfl_xor_sync(0xffffffffu, (unsigned long long)best, o);
    best = other > best ? other : best;
  }
  if ((threadIdx.x & 31) == 0 && best != 0) atomicMax(&st->pred, (unsigned long long)best);
}

template <typename K>
__global__ void select_finish_kernel(const SelectState* __restrict__ st, long long n, dfs_eer_result* __restrict__ res) {
  if (threadIdx.x != 0) return;
  typedef typename ScoreOf<K>::type S;
  const long long n_spoof = (long long)st->n_spoof, n_bona = (long long)st->n_bona;
  const long long k = st->best.idx, c1 = st->best.c1, c0 = k - c1;
  const long long start = (long long)(st->c0_below + st->c1_below);
  const double far = __ddiv_rn((double)(n_spoof - c0), (double)n_spoof);
  const double frr = __ddiv_rn((double)c1, (double)n_bona);
  res->eer = __ddiv_rn(__dadd_rn(far, frr), 2.0);                      // evaluation.py:29
  const K G = (K)st->prefix;
  const S eps = (S)1e-6;
  double thr;
  if (k == 0) thr = (double)(S)(from_key(G) - eps);                    // evaluation.py:32-33 (the smallest score is G)
  else if (k == n) thr = (double)(S)(from_key(G) + eps);               // :34-35 (the largest score is G)
  else if (k > start) thr = (double)from_key(G);                       // :37, sorted_scores[k-1] lies in G's group
  else thr = (double)from_key((K)st->pred);                            //      ... or is the largest score below G
  res->threshold = thr;
  res->eer_idx = k;
  res->n_bonafide = n_bona;
  res->n_spoof = n_spoof;
}

// 1 = cp.async.bulk ring (default), 0 = direct vector loads: kept switchable for the cross-check test (dfs_set_global_option)
int g_select_use_tma = 1;

struct SelectScratch {
  SelectState* state = nullptr;          // device
  unsigned long long* hist = nullptr;    // device [512]
  dfs_eer_result* res = nullptr;         // device
};
static SelectScratch g_sel[16];

template <typename K>
static int eer_select_impl(const void* scores_v, const uint8_t* labels, int64_t n, dfs_eer_result* result_host, cudaStream_t stream) {
  typedef typename ScoreOf<K>::type S;
  constexpr int LEVELS = sizeof(K);
  constexpr int V = VecOf<K>::V;
  const S* scores = static_cast<const S*>(scores_v);
  int dev = 0, num_sms = 148;
  DFS_CUDA_CHECK(cudaGetDevice(&dev));
  DFS_REQUIRE(dev >= 0 && dev < 16, DFS_ERR_INVALID, "device index %d out of range", dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  SelectScratch& sc = g_sel[dev];
  if (sc.state == nullptr) {
    void* p = nullptr;
    DFS_CUDA_CHECK(cudaMalloc(&p, 8192));
    sc.state = static_cast<SelectState*>(p);
    sc.hist = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(p) + 1024);
    sc.res = reinterpret_cast<dfs_eer_result*>(static_cast<uint8_t*>(p) + 1024 + 4096);
  }
  static bool configured[32] = {false};
  constexpr int kHistSmem = 512 * kCntStride;
  if (dfs_first_use_on_device(configured)) {
    DFS_CUDA_CHECK(cudaFuncSetAttribute(select_hist_kernel<uint32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmem));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(select_hist_kernel<uint32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmem));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(select_hist_kernel<uint64_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmem));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(select_hist_kernel<uint64_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmem));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(select_hist_tma_kernel<uint32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelTmaSmem));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(select_hist_tma_kernel<uint32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelTmaSmem));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(select_hist_tma_kernel<uint64_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelTmaSmem));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(select_hist_tma_kernel<uint64_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelTmaSmem));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(select_hist_scalar_kernel<uint32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmem));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(select_hist_scalar_kernel<uint32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmem));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(select_hist_scalar_kernel<uint64_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmem));
    DFS_CUDA_CHECK(cudaFuncSetAttribute(select_hist_scalar_kernel<uint64_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmem));
  }
  const bool aligned = (reinterpret_cast<uintptr_t>(scores) % 16 == 0) && (reinterpret_cast<uintptr_t>(labels) % V == 0);
  const bool tma = (reinterpret_cast<uintptr_t>(scores) % 16 == 0) && (reinterpret_cast<uintptr_t>(labels) % 16 == 0) && g_select_use_tma;
  const long long chunk = tma ? (long long)(16384 / sizeof(S)) : aligned ? 256ll * V * kSelIters : 256ll * 224;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(ceil_div64(n, chunk), num_sms));

  select_init_kernel<<<1, 256, 0, stream>>>(sc.state, sc.hist);
  DFS_LAUNCH_CHECK();
  SelectState host{};
  for (int lv = LEVELS - 1; lv >= 0; --lv) {
    const int shift = 8 * lv;
    const bool first = (lv == LEVELS - 1);
    const bool skip = !first && ((((host.key_and ^ host.key_or) >> shift) & 0xffull) == 0);
    if (!skip) {
      if (tma) {
        if (first) select_hist_tma_kernel<K, true><<<grid, 288, kSelTmaSmem, stream>>>(scores, labels, n, shift, sc.state, sc.hist);
        else select_hist_tma_kernel<K, false><<<grid, 288, kSelTmaSmem, stream>>>(scores, labels, n, shift, sc.state, sc.hist);
      } else if (aligned) {
        if (first) select_hist_kernel<K, true><<<grid, 256, kHistSmem, stream>>>(scores, labels, n, shift, sc.state, sc.hist);
        else select_hist_kernel<K, false><<<grid, 256, kHistSmem, stream>>>(scores, labels, n, shift, sc.state, sc.hist);
      } else {
        if (first) select_hist_scalar_kernel<K, true><<<grid, 256, kHistSmem, stream>>>(scores, labels, n, shift, sc.state, sc.hist);
        else select_hist_scalar_kernel<K, false><<<grid, 256, kHistSmem, stream>>>(scores, labels, n, shift, sc.state, sc.hist);
      }
      DFS_LAUNCH_CHECK();
    }
    select_pick_kernel<<<1, 256, 0, stream>>>(sc.hist, sc.state, shift, first ? 1 : 0, skip ? 1 : 0);
    DFS_LAUNCH_CHECK();
    if (first) {  // label totals (single-class early-out, evaluation.py:18-19) and the key AND / OR (constant bytes are skipped)
      DFS_CUDA_CHECK(cudaMemcpyAsync(&host, sc.state, sizeof(host), cudaMemcpyDeviceToHost, stream));
      DFS_CUDA_CHECK(cudaStreamSynchronize(stream));
      if (host.status != 0) {
        result_host->eer = 0.0;
        result_host->threshold = 0.0;
        result_host->eer_idx = -1;
        result_host->n_bonafide = (int64_t)host.n_bona;
        result_host->n_spoof = (int64_t)host.n_spoof;
        return DFS_OK;
      }
    }
  }
  select_uniform_group_kernel<<<1, 32, 0, stream>>>(sc.state);
  DFS_LAUNCH_CHECK();
  DFS_CUDA_CHECK(cudaMemcpyAsync(&host, sc.state, sizeof(host), cudaMemcpyDeviceToHost, stream));
  DFS_CUDA_CHECK(cudaStreamSynchronize(stream));
  const long long start = (long long)(host.c0_below + host.c1_below);
  if (!host.resolved) {
    // mixed-label tie group of m scores: compact its labels in index order, sweep the m + 1 curve points it spans
    const long long m = (long long)(host.g0 + host.g1);
    const long long tiles_n = ceil_div64(n, kSortTile), tiles_m = ceil_div64(m, kSortTile);
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_cnt = carve((size_t)tiles_n * 4), o_excl = carve((size_t)tiles_n * 8), o_pay = carve((size_t)m * 4);
    const size_t o_bones = carve((size_t)tiles_m * 4), o_bexcl = carve((size_t)tiles_m * 8), o_bbest = carve((size_t)tiles_m * sizeof(SweepBest));
    void* base = nullptr;
    DFS_PROPAGATE(get_workspace(off, &base));
    uint8_t* b8 = static_cast<uint8_t*>(base);
    uint32_t* tcnt = reinterpret_cast<uint32_t*>(b8 + o_cnt);
    unsigned long long* texcl = reinterpret_cast<unsigned long long*>(b8 + o_excl);
    uint32_t* gpay = reinterpret_cast<uint32_t*>(b8 + o_pay);
    uint32_t* bones = reinterpret_cast<uint32_t*>(b8 + o_bones);
    unsigned long long* bexcl = reinterpret_cast<unsigned long long*>(b8 + o_bexcl);
    SweepBest* bbest = reinterpret_cast<SweepBest*>(b8 + o_bbest);
    group_count_kernel<K><<<(unsigned)tiles_n, 256, 0, stream>>>(scores, n, sc.state, tcnt);
    DFS_LAUNCH_CHECK();
    sweep_scan_kernel<<<1, 1024, 0, stream>>>(tcnt, tiles_n, texcl, 0, 0, 0, nullptr);
    DFS_LAUNCH_CHECK();
    group_write_kernel<K><<<(unsigned)tiles_n, 256, 0, stream>>>(scores, labels, n, sc.state, texcl, gpay);
    DFS_LAUNCH_CHECK();
    sweep_count_kernel<<<(unsigned)tiles_m, 256, 0, stream>>>(gpay, m, bones);
    DFS_LAUNCH_CHECK();
    sweep_scan_kernel<<<1, 1024, 0, stream>>>(bones, tiles_m, bexcl, 0, 0, 0, nullptr);
    DFS_LAUNCH_CHECK();
    sweep_min_kernel<<<(unsigned)tiles_m, 256, 0, stream>>>(gpay, m, (long long)host.n_bona, (long long)host.n_spoof, bexcl, bones, bbest, start,
                                                           (long long)host.c1_below, /*single=*/0, nullptr);
    DFS_LAUNCH_CHECK();
    select_reduce_best_kernel<<<1, 256, 0, stream>>>(bbest, tiles_m, sc.state);
    DFS_LAUNCH_CHECK();
    DFS_CUDA_CHECK(cudaMemcpyAsync(&host, sc.state, sizeof(host), cudaMemcpyDeviceToHost, stream));
    DFS_CUDA_CHECK(cudaStreamSynchronize(stream));
  }
  if (host.best.idx == start && start > 0) {
    select_pred_kernel<K><<<(unsigned)std::max<long long>(1, std::min<long long>(ceil_div64(n, 256), (long long)num_sms * 8)), 256, 0, stream>>>(
        scores, n, sc.state);
    DFS_LAUNCH_CHECK();
  }
  select_finish_kernel<K><<<1, 32, 0, stream>>>(sc.state, n, sc.res);
  DFS_LAUNCH_CHECK();
  DFS_CUDA_CHECK(cudaMemcpyAsync(result_host, sc.res, sizeof(dfs_eer_result), cudaMemcpyDeviceToHost, stream));
  DFS_CUDA_CHECK(cudaStreamSynchronize(stream));
  return DFS_OK;
}

int eer_select_device(const void* scores, int key_bytes, const uint8_t* labels, int64_t n, dfs_eer_result* result_host, cudaStream_t stream) {
  DFS_REQUIRE(scores && labels && result_host, DFS_ERR_INVALID, "dfs_eer_select: NULL argument");
  DFS_REQUIRE(n > 0 && n < (1ll << 30), DFS_ERR_INVALID, "dfs_eer_select: n = %lld outside [1, 2^30)", (long long)n);
  DFS_REQUIRE(key_bytes == 4 || key_bytes == 8, DFS_ERR_INVALID, "dfs_eer_select: key_bytes must be 4 (fp32) or 8 (fp64)");
  return key_bytes == 4 ? eer_select_impl<uint32_t>(scores, labels, n, result_host, stream)
                        : eer_select_impl<uint64_t>(scores, labels, n, result_host, stream);
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

// ---- BCEWithLogitsLoss(reduction='mean') over a whole score vector ---------------------------------------
// src/evaluation.py:83-86 accumulates criterion(logits, labels).item() * batch per batch; here the per-element loss
//   max(x, 0) - x*y + log1p(exp(-|x|))        (torch's numerically stable form, evaluated in fp32 like torch)
// is summed once over all n scores in fp64 (fixed order per launch geometry: per-thread -> warp shuffle -> one
// atomicAdd(double) per warp is NOT order-deterministic, so the warp partials go to a buffer reduced by one thread).
__global__ void __launch_bounds__(256) bce_partial_kernel(const float* __restrict__ logits, const float* __restrict__ labels, long long n,
                                                           double* __restrict__ partial /*[gridDim.x]*/) {
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = logits[i], y = labels[i];
    const float l = fmaxf(x, 0.0f) - x * y + log1pf(expf(-fabsf(x)));
    acc += (double)l;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += part[i];
    partial[blockIdx.x] = s;
  }
}
__global__ void bce_final_kernel(const double* __restrict__ partial, int nb, long long n, double* __restrict__ out) {
  if (threadIdx.x != 0) return;
  double s = 0.0;
  for (int i = 0; i < nb; ++i) s += partial[i];
  out[0] = s / (double)n;
}

int bce_with_logits_device(const float* logits, const float* labels, int64_t n, double* mean_host, cudaStream_t stream) {
  DFS_REQUIRE(logits && labels && mean_host && n > 0, DFS_ERR_INVALID, "dfs_bce_with_logits: bad argument");
  const int nb = (int)std::min<long long>(ceil_div64(n, 256), 148 * 4);
  void* base = nullptr;
  DFS_PROPAGATE(get_workspace((size_t)(nb + 1) * 8, &base));
  double* partial = static_cast<double*>(base);
  bce_partial_kernel<<<nb, 256, 0, stream>>>(logits, labels, n, partial);
  DFS_LAUNCH_CHECK();
  bce_final_kernel<<<1, 32, 0, stream>>>(partial, nb, n, partial + nb);
  DFS_LAUNCH_CHECK();
  DFS_CUDA_CHECK(cudaMemcpyAsync(mean_host, partial + nb, 8, cudaMemcpyDeviceToHost, stream));
  DFS_CUDA_CHECK(cudaStreamSynchronize(stream));
  return DFS_OK;
}

}  // namespace dfs
