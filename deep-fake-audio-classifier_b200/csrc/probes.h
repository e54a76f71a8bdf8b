/* probes.h -- C exports of lib/libdfs_b200_probes.so: bring-up probes and micro-benchmarks (tests/test_gpu_probes.py,
 * tools/umma_bench.py, tools/micro/tmem_ld_bench.py).  Measurement code only: nothing here is part of the scoring path or of the
 * public ABI (include/dfs_b200.h), and the product library does not contain it. */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* dfs_probe_last_error(void);
/* One tcgen05.mma tile D[128,N] = A'[128,K] * B[N,K]^T.  A [rows_a,K] and B [N,K] are row-major bf16 bit patterns in device
 * memory, staged in shared memory in the layout the conv kernels use and read through the same SWIZZLE_NONE K-major
 * descriptors.  D row r = 8g+i reads staged A row (row_shift + g*group_rows + i) -- the addressing of a 3x3 tap on a
 * 16x8 tile.  out_dev [128*N] fp32. */
int dfs_probe_umma(const uint16_t* a_dev, const uint16_t* b_dev, int rows_a, int n, int k, int row_shift, int group_rows, float* out_dev,
                   void* stream);
/* One 3-D TMA box load (wrows rows x 18 columns x all planes) from an FT8 activation buffer at (row0, col0); the
 * shared-memory image is copied to out_dev [planes*18*wrows*8] bf16 bits. */
int dfs_probe_tma_window(const uint16_t* act_dev, int planes, int rs, int64_t ncols, int wrows, int row0, int col0, uint16_t* out_dev,
                         void* stream);
/* Issue-rate / operand-fetch micro-benchmark of tcgen05.mma (M=128, N=n, K=16): `iters` rounds of `nmma` (<= 96) MMAs whose
 * A/B descriptor start addresses are smem_base + a_off[i] / b_off[i] (bytes), rotating over `n_acc` accumulators (1 = one
 * dependent chain), one commit + wait per round; cycles_host receives the SM cycles of the timed rounds.  Bit 1 of
 * use_base_offset runs the loop on a CTA pair (cta_group::2). */
int dfs_probe_umma_bench(int n, int nmma, int iters, int n_acc, const uint32_t* a_off_host, const uint32_t* b_off_host, uint32_t a_lbo,
                         uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, uint32_t layout, uint32_t use_base_offset, int64_t* cycles_host,
                         void* stream);
/* TMEM read-out rate: `nwarps` (4 | 8 | 16) warps of each of `blocks` CTAs read their lane quadrant with one tcgen05.ld shape
 * (0..4 = 32x32b .x8 .x16 .x32 .x64 .x128, 5..7 = 16x256b .x4 .x8 .x16, 8..10 = 16x128b .x8 .x16 .x32), `lds_per_wait` loads
 * per tcgen05.wait::ld, `iters` rounds.  Returns the slowest block's SM cycles and the bytes one block moved. */
int dfs_probe_tmem_ld_bench(int shape, int nwarps, int blocks, int iters, int lds_per_wait, int64_t* cycles_host,
                            int64_t* bytes_per_block_host, void* stream);

#ifdef __cplusplus
}
#endif
