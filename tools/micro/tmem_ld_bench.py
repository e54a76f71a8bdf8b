"""TMEM read-out rate (tcgen05.ld) per SM on this GPU, by shape, warps per CTA and loads in flight.
Run on the GPU box:  python tools/micro/tmem_ld_bench.py > gpurun_out/tmem_ld_bench.txt

Why: conv1 / CAE enc1 read 14.7 MB of fp32 accumulators per utterance out of TMEM; round 1 quoted "64 B/clk/SM" for that
read-out without a measurement behind it.  This prints bytes/clk/SM for one CTA and for 148 CTAs."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "deep-fake-audio-classifier_b200"))
import torch  # noqa: E402

from dfs_b200 import _probes as P  # noqa: E402

torch.zeros(1, device="cuda")
lib = P.load()
SHAPES = ["32x32b.x8", "32x32b.x16", "32x32b.x32", "32x32b.x64", "32x32b.x128", "16x256b.x4", "16x256b.x8", "16x256b.x16",
          "16x128b.x8", "16x128b.x16", "16x128b.x32"]


def rate(shape, nwarps, blocks, lpw, iters=2000):
    cyc, byt = C.c_int64(), C.c_int64()
    P.check(lib.dfs_probe_tmem_ld_bench(shape, nwarps, blocks, iters, lpw, C.byref(cyc), C.byref(byt), None), "tmem_ld_bench")
    return byt.value / cyc.value


print("# tcgen05.ld read-out: bytes / clk / SM (one CTA per SM; lpw = loads issued before each tcgen05.wait::ld)")
print(f"{'shape':12s} {'warps':>5s} {'lpw':>3s} {'1 CTA':>9s} {'148 CTAs':>9s}")
for s, name in enumerate(SHAPES):
    for nwarps in (4, 8, 16):
        if s == 4 and nwarps == 16:
            continue
        for lpw in (1, 2, 4):
            print(f"{name:12s} {nwarps:5d} {lpw:3d} {rate(s, nwarps, 1, lpw):9.1f} {rate(s, nwarps, 148, lpw):9.1f}")
