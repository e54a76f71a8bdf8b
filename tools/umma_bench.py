"""tcgen05.mma cost model on this GPU: cycles per MMA for the addressing patterns the conv kernels use.
Run on the GPU box:  python tools/umma_bench.py > gpurun_out/umma_bench.txt"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-fake-audio-classifier_b200"))
import torch  # noqa: E402

from dfs_b200 import _probes as N  # noqa: E402

torch.zeros(1, device="cuda")
lib = N.load()


def cycles(n, a_off, b_off, a_lbo, a_sbo, b_lbo, b_sbo, layout=0, base=0, iters=200, n_acc=1):
    nm = len(a_off)
    A = (C.c_uint32 * nm)(*a_off)
    B = (C.c_uint32 * nm)(*b_off)
    cyc = C.c_int64()
    N.check(lib.dfs_probe_umma_bench(n, nm, iters, n_acc, A, B, a_lbo, a_sbo, b_lbo, b_sbo, layout, base, C.byref(cyc), None), "umma_bench")
    return cyc.value / iters


def run(name, n, a_off, b_off, a_lbo, a_sbo, b_lbo, b_sbo, layout=0, base=0, iters=200, n_acc=1):
    per = cycles(n, a_off, b_off, a_lbo, a_sbo, b_lbo, b_sbo, layout, base, iters, n_acc) / len(a_off)
    print(f"{name:58s} N={n:3d} nmma={len(a_off):2d} acc={n_acc}  {per:7.1f} cyc/MMA   ideal {128 * n / 256:5.1f}")


def sweep(name, n, b_lbo, accs=(1, 2, 4), nmmas=(8, 16, 32, 64, 96), base=0):
    """Rounds of nmma MMAs with one commit + wait per round: cycles(round) = fixed + nmma * per_mma.  The slope between the two
    longest rounds is the steady-state cost of one MMA, the intercept the commit / wait / pipeline-fill latency that a
    per-round average (what round 1 reported) smears over the MMAs."""
    for n_acc in accs:
        if n_acc * n > 512:
            continue
        pts = [(m, cycles(n, [0] * m, [0] * m, 18 * 8 * 16, 128, b_lbo, 128, n_acc=n_acc, base=base)) for m in nmmas]
        (m0, c0), (m1, c1) = pts[-2], pts[-1]
        slope = (c1 - c0) / (m1 - m0)
        print(f"{name:40s} N={n:3d} acc={n_acc}  " + "  ".join(f"{m}:{c / m:6.1f}" for m, c in pts) +
              f"   slope {slope:6.1f} cyc/MMA  fixed {c1 - slope * m1:6.0f} cyc   floor {128 * n / 256:5.1f}")


print("# nmma sweep (cycles per MMA averaged over a round of nmma; slope = steady-state cost, fixed = per-round latency)")
print("# lean issue loop: fixed descriptors in registers, nothing between two tcgen05.mma")
for n_, lbo_ in ((32, 512), (64, 1024), (128, 2048), (256, 4096)):
    sweep("lean  none aligned, fixed A/B", n_, lbo_, base=4)
print("# table-driven issue loop (descriptors loaded from shared memory per MMA, as in round 1)")
for n_, lbo_ in ((64, 1024), (128, 2048), (256, 4096)):
    sweep("table none aligned, fixed A/B", n_, lbo_)


# --- no swizzle, conv3-like: CIN=64 (4 k-steps), WROWS=10, plane = 18*10*16 = 2880 B
P3 = 2880
taps3 = [(2 * kk) * P3 + (kw * 10 + kh) * 16 for kh in range(3) for kw in range(3) for kk in range(4)]
b3 = [((t * 8 + 2 * kk) * 128) * 16 % 65536 for t in range(9) for kk in range(4)]
run("none  conv3 taps (SBO=160, 16B-granular shifts)", 128, taps3, b3, P3, 160, 2048, 128)
al3 = [(2 * kk) * P3 for _ in range(9) for kk in range(4)]
run("none  conv3 K-steps only, start aligned, SBO=160", 128, al3, b3, P3, 160, 2048, 128)
P3a = 18 * 8 * 16
run("none  aligned core matrices (SBO=128, 128B shifts)", 128, [(2 * kk) * P3a + kw * 128 for kh in range(3) for kw in range(3) for kk in range(4)], b3, P3a, 128, 2048, 128)
run("none  aligned, fixed A (same address every MMA)", 128, [0] * 36, [0] * 36, P3a, 128, 2048, 128)
run("none  aligned, N=256", 256, [0] * 16, [0] * 16, P3a, 128, 4096, 128)
run("none  aligned, N=64", 64, [0] * 18, [0] * 18, P3a, 128, 1024, 128)
# --- conv2-like: CIN=32 (2 k-steps), WROWS=18, plane = 18*18*16 = 5184
P2 = 5184
taps2 = [(2 * kk) * P2 + (kw * 18 + kh) * 16 for kh in range(3) for kw in range(3) for kk in range(2)]
b2 = [((t * 4 + 2 * kk) * 64) * 16 for t in range(9) for kk in range(2)]
run("none  conv2 taps (SBO=288), N=64", 64, taps2, b2, P2, 288, 1024, 128)
run("none  conv2 taps (SBO=288), N=128 (as if 2 tiles shared B)", 128, taps2, b2, P2, 288, 2048, 128)
# --- 128B swizzle, K-major: rows of 128 B (64 ch); group of 8 rows = 1024 B
sw_al = [kk * 32 for _ in range(9) for kk in range(4)]
bsw = [(t * 16384 + kk * 32) % 65536 for t in range(9) for kk in range(4)]
run("SW128 aligned tile (SBO=1024), K advance by +32 B", 128, sw_al, bsw, 16, 1024, 16, 1024, layout=2)
sw_taps = [(kw * 10 + kh) * 128 + kk * 32 for kh in range(3) for kw in range(3) for kk in range(4)]
run("SW128 conv3 taps: row shifts of 128 B, SBO=1280", 128, sw_taps, bsw, 16, 1280, 16, 1024, layout=2)
run("SW128 conv3 taps + base_offset field", 128, sw_taps, bsw, 16, 1280, 16, 1024, layout=2, base=1)
sw_taps8 = [(kw * 8) * 128 + kh * 128 + kk * 32 for kh in range(3) for kw in range(3) for kk in range(4)]
run("SW128 taps, window rows = 8 (SBO=1024)", 128, sw_taps8, bsw, 16, 1024, 16, 1024, layout=2)
# --- 64B swizzle (CIN=32 rows of 64 B)
sw64 = [(kw * 18 + kh) * 64 + kk * 32 for kh in range(3) for kw in range(3) for kk in range(2)]
b64 = [(t * 4096 + kk * 32) for t in range(9) for kk in range(2)]
run("SW64  conv2 taps: row shifts of 64 B, SBO=18*64, N=64", 64, sw64, b64, 16, 18 * 64, 16, 512, layout=4)
run("SW64  aligned (SBO=512), N=64", 64, [kk * 32 for _ in range(9) for kk in range(2)], b64, 16, 512, 16, 512, layout=4)

# --- CTA pairs (cta_group::2, M = 256): bit 1 of the last flag.  B strides are those of the per-CTA half image (n/2 rows).
print("# cta_group::2 (M = 256 per MMA; 'ideal' is per CTA pair = 128 * n / 256 cycles)")
for n in (64, 128, 256):
    run(f"pair  none aligned, N={n} (B half = {n // 2} rows per CTA)", n, [0] * 16, [0] * 16, P3a, 128, (n // 2) * 16, 128, base=2)
run("pair  none conv3-like taps, N=128", 128, taps3, [((t * 8 + 2 * kk) * 64) * 16 % 65536 for t in range(9) for kk in range(4)], P3, 160, 1024, 128, base=2)
run("pair  SW128 aligned, N=128", 128, sw_al, [(t * 8192 + kk * 32) % 65536 for t in range(9) for kk in range(4)], 16, 1024, 16, 1024, layout=2, base=2)
run("pair  SW128 aligned, N=256", 256, sw_al[:16], [(kk * 32) for _ in range(4) for kk in range(4)], 16, 1024, 16, 1024, layout=2, base=2)
