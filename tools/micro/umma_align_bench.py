"""Does an SS-mode tcgen05.mma pay for A core matrices that are not 128-byte aligned?  Lean issue loop (tools/umma_bench.py), slope of an nmma
sweep = steady-state cycles per MMA, for the A addressing of the conv kernels: SBO = 160 B (10-row windows), 288 B (18-row windows),
start addresses shifted by 16-byte rows (the taps), against 128-byte aligned core matrices."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "deep-fake-audio-classifier_b200"))
import torch  # noqa: E402

from dfs_b200 import _probes as N  # noqa: E402

torch.zeros(1, device="cuda")
lib = N.load()


def cycles(n, nm, a_off, a_lbo, a_sbo, b_lbo, iters=200, n_acc=1):
    A = (C.c_uint32 * nm)(*([a_off] * nm))
    B = (C.c_uint32 * nm)(*([0] * nm))
    cyc = C.c_int64()
    N.check(lib.dfs_probe_umma_bench(n, nm, iters, n_acc, A, B, a_lbo, a_sbo, b_lbo, 128, 0, 4, C.byref(cyc), None), "umma_bench")
    return cyc.value / iters


def slope(n, a_off, a_lbo, a_sbo, b_lbo):
    c0, c1 = cycles(n, 64, a_off, a_lbo, a_sbo, b_lbo), cycles(n, 96, a_off, a_lbo, a_sbo, b_lbo)
    return (c1 - c0) / 32


print("# cycles per MMA (M = 128, K = 16, kind::f16, SS mode), lean issue loop, slope between rounds of 64 and 96 MMAs")
print(f"{'A addressing':58s}" + "".join(f"  N={n:3d}" for n in (32, 64, 128, 256)))
for name, off, lbo, sbo in (("aligned core matrices: SBO 128, start +0", 0, 2304, 128),
                            ("SBO 128, start +16 (one row down)", 16, 2304, 128),
                            ("SBO 128, start +64", 64, 2304, 128),
                            ("SBO 160 (10-row window), start +0", 0, 2880, 160),
                            ("SBO 160, start +16", 16, 2880, 160),
                            ("SBO 160, start +32", 32, 2880, 160),
                            ("SBO 288 (18-row window), start +0", 0, 5184, 288),
                            ("SBO 288, start +16", 16, 5184, 288),
                            ("SBO 256 (16-row window, aligned), start +0", 0, 4608, 256),
                            ("SBO 256, start +16", 16, 4608, 256),
                            ("SBO 16*41 = 656 (Toeplitz conv1 / enc1), LBO 16, start +0", 0, 16, 656),
                            ("SBO 128, LBO 16 (Toeplitz, conv1_tc), start +0", 0, 16, 128)):
    print(f"{name:58s}" + "".join(f"  {slope(n, off, lbo, sbo, n * 16):5.1f}" for n in (32, 64, 128, 256)), flush=True)

print("# A operand: K-chunk stride (LBO) modulo 128 bytes, SBO 160, start +16")
print(f"{'LBO':58s}" + "".join(f"  N={n:3d}" for n in (32, 64, 128, 256)))
for extra in (0, 16, 32, 48, 64, 80, 96, 112):
    lbo = 2560 + extra
    print(f"{'LBO = 2560 + %3d' % extra:58s}" + "".join(f"  {slope(n, 16, lbo, 160, n * 16):5.1f}" for n in (32, 64, 128, 256)), flush=True)
print("# B operand: K-chunk stride (LBO) = rows x 16 B (+ pad), N rows")
for pad in (0, 16, 64):
    print(f"{'B LBO = N * 16 + %2d' % pad:58s}" + "".join(f"  {slope(n, 16, 2880, 160, n * 16 + pad):5.1f}" for n in (32, 64, 128, 256)), flush=True)
