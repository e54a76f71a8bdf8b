"""Build libdfs_b200.so in-tree with nvcc for sm_100a (no torch dependency in the library).

    python deep-fake-audio-classifier_b200/build.py [--force]

Objects are cached by source mtime under build/ ; the shared library lands in lib/ (git-ignored,
travels to the GPU box with the gpurun snapshot).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", "libdfs_b200.so")
OUT_PROBES = os.path.join(HERE, "lib", "libdfs_b200_probes.so")   # bring-up probes + micro-benchmarks (tests / tools only)
OBJ = os.path.join(HERE, "build")
SOURCES = ["api.cu", "conv_tc.cu", "conv1_tc.cu", "conv12_fused.cu", "cae_tc.cu", "cae_enc1_tc.cu", "cnn1d_tc.cu", "cnn1d_l1_fused.cu", "cnn1d_fused.cu", "cnn2d.cu", "cnn2d_fp32.cu", "simt_models.cu", "eer.cu", "synth.cu"]
PROBE_SOURCES = ["probe.cu", "conv_tc.cu"]   # probe.cu uses conv_tc.cu's tensor-map helper
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "dfs_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force, hdr_m):
    s = os.path.join(CSRC, src)
    o = os.path.join(OBJ, src.replace(".cu", ".o"))
    if not force and os.path.exists(o) and os.path.getmtime(o) > max(os.path.getmtime(s), hdr_m):
        return o, ""
    r = subprocess.run([NVCC, *FLAGS, "-c", s, "-o", o], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return o, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    hdr_m = _deps_mtime()
    all_sources = SOURCES + [s for s in PROBE_SOURCES if s not in SOURCES]
    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(lambda s: _compile(s, force, hdr_m), all_sources))
    by_src = {s: o for s, (o, _) in zip(all_sources, results)}
    objs = [by_src[s] for s in SOURCES]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if force or not os.path.exists(OUT) or any(os.path.getmtime(o) > os.path.getmtime(OUT) for o in objs):
        _link(OUT, objs)
    pobjs = [by_src[s] for s in PROBE_SOURCES]
    if force or not os.path.exists(OUT_PROBES) or any(os.path.getmtime(o) > os.path.getmtime(OUT_PROBES) for o in pobjs):
        _link(OUT_PROBES, pobjs)
    return OUT


def _link(out, objs):
    # --cudart shared: the artefact binds to libcudart.so at load time (torch ships one) instead of carrying a static copy of the
    # whole runtime, entry points this code never calls included; -Bsymbolic: each library resolves its own helpers internally
    tmp = out + ".tmp"
    r = subprocess.run([NVCC, "-shared", "--cudart", "shared", "-Xlinker", "-Bsymbolic", "-o", tmp, *objs, "-lcuda",
                        "-gencode", "arch=compute_100a,code=sm_100a"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, out)    # atomic: a reader (or a snapshot of the tree) never sees a half-written library


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--force" in sys.argv))
