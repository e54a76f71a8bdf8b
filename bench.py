#!/usr/bin/env python
"""bench.py -- utterances/sec of the scoring hot path (BASELINE.json metric, configs[1]):
2D-CNN batch scoring of >= 1M synthetic [321x180] utterances, fp16 operands / fp32 accumulation,
sharded over N GPUs (one process per GPU) with one NCCL all-gather of the scores per step and the
EER of the gathered scores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass over the rank's resident pool of P utterances (default 50,000 = 11.6 GB of fp32
features, far larger than the 126 MB L2; 20 steps = 1.0 M utterances per GPU) through
conv1 -> conv2 -> conv3 -> head (+ all-gather at N > 1) + the EER of the step's scores.
Rank 0 prints ONE JSON line (keys: the driver contract, DESIGN.md "Measurement").  Besides the headline
the same line carries, under "workloads", short legs for the other BASELINE configs -- CAE-MSE (3), the
hybrid ensemble (4; device-resident and from host memory through ONE upload), the EER of 100 M scores
(5; sort and select) and the 1D-CNN alone -- each with its own roofline and a clock record taken under
its own load.  `--workload X` prints one of those legs as a line of its own (profiling runs).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "deep-fake-audio-classifier_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

FLOP_PER_UTT = {"cnn2d": 3_218_376_960, "cae": 1_792_021_760, "cnn1d": 30_816_256}
CONV3_FLOP_PER_UTT = 2 * 1_061_683_200
CONV2_FLOP_PER_UTT = 2 * 530_841_600
BYTES_PER_UTT = 321 * 180 * 4
NOMINAL_TFLOPS = 2250.0   # B200 dense bf16 / fp16 datasheet figure (B200_PROFILING.md), for context next to the measured peaks
METRIC = "utterances/sec scoring [321x180] LFCC maps (2D-CNN) + EER"
WORKLOAD = ("BASELINE configs[1]: 2D-CNN (src/model.py) batch scoring of synthetic [321x180] utterances + EER per step "
            "(ours: fp16 tensor-core operands / fp32 accumulate; reference arm: torch CPU fp32, predict.py loop, bs 32)")
SIDE = ("cae", "cnn1d", "hybrid", "eer")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pool", type=int, default=50000, help="utterances resident per GPU and scored per step (x 20 steps = 1 M)")
    ap.add_argument("--chunk", type=int, default=0, help="utterances per internal pass (0 = library default 416)")
    ap.add_argument("--e2e-pool", type=int, default=16640, help="utterances in pinned host memory for the e2e legs (one step = one call over all of them)")
    ap.add_argument("--e2e-seconds", type=float, default=1.2, help="minimum timed duration of the e2e leg (and >= 10 steps)")
    ap.add_argument("--leg-seconds", type=float, default=1.2, help="minimum timed duration of each side-workload leg")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="headline only: skip the `workloads` legs")
    ap.add_argument("--workload", default="cnn2d", choices=["cnn2d", *SIDE],
                    help="cnn2d = the headline BASELINE configs[1] (with the other configs as `workloads` legs); cae / hybrid / eer = configs "
                         "3 / 4 / 5, cnn1d = the 1D-CNN alone, each as a line of its own")
    ap.add_argument("--eer-n", type=int, default=100_000_000)
    ap.add_argument("--eer-method", default="sort", choices=["sort", "select"],
                    help="eer workload: 'sort' = full stable radix sort + sweep (north_star wording; `value`), 'select' = radix select of "
                         "the crossing (what calculate_eer() uses when no permutation is requested); the other one is timed as an extra key")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops_burst=p["bf16_tflops"], tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm_gbs=p["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock, power and throttle reasons of ONE GPU sampled every 25 ms for the whole run, on a thread of this process,
    through NVML (the counters `nvidia-smi --query-gpu=clocks.sm,...` prints; a fresh nvidia-smi process needs ~1 s before its
    first sample, longer than some legs).  Falls back to an `nvidia-smi -lms` child if pynvml is missing.  `window(t0, t1)`
    summarises the samples taken while a leg ran, so that every number carries a clock record taken under ITS load."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, torch, index):
        self.rows, self.stop_flag, self.thread, self.proc, self.source = [], False, None, None, None
        self.index = index
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
        except Exception:
            self.nv = None

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nv, self.handle
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                watts = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.rows.append((time.time(), mhz, self.max_mhz, watts, mask))
            except Exception:
                pass
            time.sleep(0.025)

    def _read(self):
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            try:
                mask = sum(bit for (_, bit), v in zip(self.REASONS, c[3:7]) if v.lower().startswith("active"))
                self.rows.append((time.time(), float(c[0]), float(c[1]), float(c[2]), mask))
            except Exception:
                pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()

    def window(self, t0, t1):
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        sm = sorted(r[1] for r in rows)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[2] for r in rows), default=None),
                "reasons": [n for n, bit in self.REASONS if any(r[4] & bit for r in rows)],
                "power_w_max": max((r[3] for r in rows), default=None), "samples": len(rows), "source": self.source}


class Ctx:
    """What every leg needs: torch, the device, rank / world, the clock sampler, the measured peaks."""

    def __init__(self, args, torch, dist, rank, world, local):
        self.args, self.torch, self.dist, self.rank, self.world, self.local = args, torch, dist, rank, world, local
        self.dev = torch.device("cuda", local)
        self.pk = peaks()
        self.sampler = ClockSampler(torch, local)
        self.sampler.start()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, step, steps, warmup=3, after_step=None):
        """`warmup` untimed steps, then EXACTLY `steps` steps between barrier + synchronize, CUDA events on the launching
        stream, max over ranks.  Returns (ms total, clock record of the timed window, last step result)."""
        torch = self.torch
        res = None
        for _ in range(warmup):
            res = step()
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        ev0.record()
        for _ in range(steps):
            res = step()
            if after_step is not None:
                after_step()
        ev1.record()
        self.barrier()
        t1 = time.time()
        return self.max_over_ranks(ev0.elapsed_time(ev1)), self.sampler.window(t0, t1), res

    def timed_for(self, step, seconds, min_steps=3, warmup=3):
        """Like timed(), with the step count chosen so that the timed region lasts >= `seconds` (same count on every rank)."""
        torch = self.torch
        for _ in range(warmup):
            step()
        self.barrier()
        t0 = time.perf_counter()
        step()
        torch.cuda.synchronize()
        one = self.max_over_ranks(time.perf_counter() - t0)
        steps = int(max(min_steps, math.ceil(seconds / max(one, 1e-6))))
        ms, clocks, res = self.timed(step, steps, warmup=0)
        return ms, steps, clocks, res


def cpu_reference_rate(torch, feats_cpu, sd, seconds, all_threads=True):
    """Times the oracle's restatement of the predict.py loop (bs 32, no_grad, all host threads) on a bounded sample."""
    from oracle import models_torch as ot
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores if all_threads else 1)
    ot.reference_loop_supervised(ot.cnn2d_forward, sd, feats_cpu[:8])          # warm the oneDNN primitives
    t0 = time.perf_counter()
    probe = ot.reference_loop_supervised(ot.cnn2d_forward, sd, feats_cpu[:32])
    rate = 32 / (time.perf_counter() - t0)
    n = int(min(feats_cpu.shape[0], max(32, (rate * seconds) // 32 * 32)))
    t0 = time.perf_counter()
    scores = ot.reference_loop_supervised(ot.cnn2d_forward, sd, feats_cpu[:n])
    dt = time.perf_counter() - t0
    del probe
    return n / dt, n, cores, scores


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle port of the predict.py loop,
    /root/reference is pure Python and is not installable as a package) on the box's host cores; rank 0 only."""
    if rank != 0:
        return
    import torch

    from dfs_b200 import synthetic as syn
    from oracle import eer as oeer
    from oracle import models_torch as ot
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = syn.cnn2d_state(0)
    pool = torch.from_numpy(syn.features(256, seed=1234))
    ot.reference_loop_supervised(ot.cnn2d_forward, sd, pool[:8])
    t0 = time.perf_counter()
    ot.reference_loop_supervised(ot.cnn2d_forward, sd, pool[:32])
    rate = 32 / (time.perf_counter() - t0)
    per_step = int(max(32, min(256, (rate * 120.0 / (args.steps + args.warmup)) // 32 * 32)))
    lab = syn.labels(per_step)
    for _ in range(args.warmup):
        ot.reference_loop_supervised(ot.cnn2d_forward, sd, pool[:per_step])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s = ot.reference_loop_supervised(ot.cnn2d_forward, sd, pool[:per_step])
        oeer.calculate_eer(s, lab)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "utterances/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "utterances_per_step": per_step, "sample": "bounded sample of the same workload sized for a few-minute run"},
        "cpu_baseline": {"value": value, "unit": "utterances/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} utterances/step x {args.steps} steps, torch CPU fp32, {cores} threads"},
        "e2e": {"value": value, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------------------------------
# raw host -> device ceiling at N concurrent ranks
# ---------------------------------------------------------------------------------------------------------------------
def h2d_ceiling(ctx, host_pool, seconds=0.5):
    """Aggregate pinned-host -> device copy rate with all ranks copying at once: bare cudaMemcpyAsync of the e2e pool in the
    2D-CNN's pass-sized pieces (416 utterances = 96 MB), no kernels.  What the e2e number can reach at most on this host."""
    torch = ctx.torch
    piece = 416
    n = host_pool.shape[0]
    dst = torch.empty((2, piece) + tuple(host_pool.shape[1:]), dtype=host_pool.dtype, device=ctx.dev)

    def sweep():
        k = 0
        for i in range(0, n, piece):
            m = min(piece, n - i)
            dst[k & 1, :m].copy_(host_pool[i:i + m], non_blocking=True)
            k += 1

    ms, steps, _, _ = ctx.timed_for(sweep, seconds, min_steps=3, warmup=1)
    nbytes = n * host_pool[0].numel() * host_pool.element_size()
    return nbytes * steps * ctx.world / (ms * 1e-3) / 1e9      # GB/s, all ranks together


# ---------------------------------------------------------------------------------------------------------------------
# side workloads (BASELINE configs 3, 4, 5 and the 1D-CNN alone)
# ---------------------------------------------------------------------------------------------------------------------
class Side:
    """Scorers and inputs shared by the side legs; built once."""

    def __init__(self, ctx, pool, labels_global, c2=None, host_pool=None):
        import dfs_b200 as D
        from dfs_b200 import synthetic as syn
        self.ctx, self.D, self.syn = ctx, D, syn
        self.pool, self.labels_global, self.host_pool = pool, labels_global, host_pool
        self._c2, self._c1, self._cae = c2, None, None

    @property
    def c2(self):
        if self._c2 is None:
            self._c2 = self.D.Cnn2dScorer(self.syn.cnn2d_state(0), device=self.ctx.local, max_chunk=self.ctx.args.chunk)
        return self._c2

    @property
    def c1(self):
        if self._c1 is None:
            self._c1 = self.D.Cnn1dScorer(self.syn.cnn1d_state(0), device=self.ctx.local)
        return self._c1

    @property
    def cae(self):
        if self._cae is None:
            mean, std = self.syn.normalizer_stats(1)
            self._cae = self.D.CaeScorer(self.syn.cae_state(0), mean, std, device=self.ctx.local)
            for key, env in (("pair_mma", "DFS_BENCH_PAIR_MMA"), ("enc3_swap", "DFS_BENCH_ENC3_SWAP"), ("dec_wide", "DFS_BENCH_DEC_WIDE")):
                if os.environ.get(env):                         # A/B switches for the kernel variants
                    self._cae.set_option(key, int(os.environ[env]))
        return self._cae


def _tensor_roof(ctx, value_per_gpu, flop, note):
    ach = value_per_gpu * flop / 1e12
    pk = ctx.pk
    return {"bound": "tensor", "achieved": ach, "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tflops_sustained"],
            "frac_of_burst_peak": ach / pk["tflops_burst"], "frac_of_nominal_peak": ach / NOMINAL_TFLOPS, "traffic": None,
            "peak_source": pk["source"] + ", sustained (timed inside a long step)", "note": note}


def _hbm_roof(ctx, gbs, note):
    pk = ctx.pk
    return {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"], "traffic": None,
            "peak_source": pk["source"], "note": note}


def leg_cae(ctx, side, seconds):
    from dfs_b200.distributed import gather_scores
    D, P, W = side.D, side.pool.shape[0], ctx.world
    n = min(P, 28 * 592)                                         # whole passes of 592; 3.8 GB of fp32 features, >> L2
    x, lab = side.pool[:n], side.labels_global[:n * W]
    cae = side.cae

    def step():
        return D.eer_details(gather_scores(cae.score(x), n_total=n * W), lab)

    ms, steps, clocks, res = ctx.timed_for(step, seconds)
    value = n * W * steps / (ms * 1e-3)
    return {"metric": "utterances/sec CAE reconstruction-MSE scoring [321x180] + EER (BASELINE configs[2])", "value": value,
            "unit": "utterances/s", "ms_per_step": ms / steps, "steps": steps, "utterances_per_step_per_gpu": n, "dtype": "f16",
            "roofline": _tensor_roof(ctx, value / W, FLOP_PER_UTT["cae"], "whole-path algorithmic FLOPs per utterance (1,792,021,760) / device time"),
            "eer": res["eer"], "clocks": clocks}


def leg_cnn1d(ctx, side, seconds):
    from dfs_b200.distributed import gather_scores
    D, n, W = side.D, side.pool.shape[0], ctx.world
    c1 = side.c1

    def step():
        return D.eer_details(gather_scores(c1.score(side.pool, apply_sigmoid=True), n_total=n * W), side.labels_global)

    ms, steps, clocks, res = ctx.timed_for(step, seconds)
    value = n * W * steps / (ms * 1e-3)
    return {"metric": "utterances/sec 1D-CNN scoring [321x180] + EER", "value": value, "unit": "utterances/s", "ms_per_step": ms / steps,
            "steps": steps, "utterances_per_step_per_gpu": n, "dtype": "f16",
            "roofline": _hbm_roof(ctx, value / W * BYTES_PER_UTT / 1e9, "231,120 B/utterance algorithmic (the fp32 input read, SURVEY.md 8d)"),
            "eer": res["eer"], "clocks": clocks}


def leg_hybrid(ctx, side, seconds):
    """BASELINE configs[3]: 2D-CNN + 1D-CNN + CAE-MSE on the same utterances, ensemble mean (src/ensemble.py:121), alpha blend
    (src/predict_hybrid.py:149-151), EER.  Device-resident, and end to end from pinned host memory through ONE upload
    (dfs_group_score_host) next to the three separate uploads the reference's per-model loops correspond to."""
    from dfs_b200.distributed import gather_scores
    D, P, W, torch = side.D, side.pool.shape[0], ctx.world, ctx.torch
    n = min(P, 7 * 2368)                                         # whole group slabs (2,368 = 4 x 592); 3.8 GB, >> L2
    x, lab = side.pool[:n], side.labels_global[:n * W]
    c2, c1, cae = side.c2, side.c1, side.cae

    def combine(s2, s1, mse):
        g2, g1, gm = (gather_scores(v, n_total=v.numel() * W) for v in (s2, s1, mse))
        sup = D.ensemble_mean([g2, g1], as_numpy=False)
        return D.eer_details(D.hybrid_blend(sup, gm, 0.8, as_numpy=False), side.labels_global[:g2.numel()])

    def step():
        return combine(c2.score(x, apply_sigmoid=True), c1.score(x, apply_sigmoid=True), cae.score(x))

    ms, steps, clocks, res = ctx.timed_for(step, seconds)
    value = n * W * steps / (ms * 1e-3)
    out = {"metric": "utterances/sec hybrid scoring (2D-CNN + 1D-CNN + CAE-MSE, blend alpha=0.8) [321x180] + EER (BASELINE configs[3])",
           "value": value, "unit": "utterances/s", "ms_per_step": ms / steps, "steps": steps, "utterances_per_step_per_gpu": n, "dtype": "f16",
           "roofline": _tensor_roof(ctx, value / W, sum(FLOP_PER_UTT.values()), "algorithmic FLOPs of the three models per utterance (5,041,214,976) / device time"),
           "eer": res["eer"], "clocks": clocks}
    if side.host_pool is not None:
        hp = side.host_pool
        ne = hp.shape[0]
        group = D.ScorerGroup([c2, c1, cae])

        def as_dev(v):
            return torch.from_numpy(v).to(ctx.dev)

        def e2e_once():
            s2, s1, mse = group.score_host(hp)
            return combine(as_dev(s2), as_dev(s1), as_dev(mse))

        def e2e_thrice():
            return combine(as_dev(c2.score_host(hp, 1)), as_dev(c1.score_host(hp, 1)), as_dev(cae.score_host(hp)))

        same = all((a == b).all() for a, b in zip(group.score_host(hp), (c2.score_host(hp, 1), c1.score_host(hp, 1), cae.score_host(hp))))
        ms1, st1, ck1, _ = ctx.timed_for(e2e_once, seconds, min_steps=5, warmup=1)
        ms3, st3, _, _ = ctx.timed_for(e2e_thrice, seconds / 2, min_steps=3, warmup=1)
        v1, v3 = ne * W * st1 / (ms1 * 1e-3), ne * W * st3 / (ms3 * 1e-3)
        out["e2e"] = {"value": v1, "unit": "utterances/s", "h2d_bytes_per_step": ne * BYTES_PER_UTT, "d2h_bytes_per_step": 3 * ne * 4,
                      "utterances_per_step_per_gpu": ne, "steps": st1, "seconds": ms1 * 1e-3, "clocks": ck1,
                      "frac_of_device_resident": v1 / value, "scores_identical_to_separate_calls": bool(same),
                      "three_uploads": {"value": v3, "h2d_bytes_per_step": 3 * ne * BYTES_PER_UTT, "steps": st3},
                      "note": "dfs_group_score_host: every slab of the pinned table is uploaded ONCE and scored by all three models; "
                              "three_uploads = one dfs_score_host call per model (what the reference's per-model passes amount to)"}
        group.close()
    return out


def eer_inputs(ctx, n, which, cache):
    """Score / label vectors of the EER leg, built on the device.  'affine' = the tie-free LCG stride pattern of
    dfs_b200.synthetic.tie_free_scores (distinct fp32 bit patterns), 'permuted' = the same values under a random permutation
    (no arithmetic structure in the index -> key map), 'sigmoid' = sigmoid of N(0, 4^2) logits in fp32: heavy ties at 0 / 1."""
    torch = ctx.torch
    from dfs_b200 import synthetic as syn
    g = torch.Generator(device=ctx.dev)
    g.manual_seed(1234)
    if which in ("affine", "permuted"):
        if "affine" not in cache:
            sc, lab = syn.tie_free_scores(n, seed=6)
            cache["affine"] = (torch.from_numpy(sc).to(ctx.dev), torch.from_numpy(lab).to(ctx.dev))
        s, l = cache["affine"]
        if which == "permuted":
            perm = torch.randperm(n, device=ctx.dev, generator=g)
            s, l = s[perm].contiguous(), l[perm].contiguous()
        return s, l
    cache.clear()
    logit = 4.0 * torch.randn(n, device=ctx.dev, generator=g)
    s = torch.sigmoid(logit)
    l = (torch.rand(n, device=ctx.dev, generator=g) < torch.sigmoid(0.5 * logit)).to(torch.uint8)
    return s.contiguous(), l.contiguous()


def leg_eer(ctx, side_D, seconds, n, method="sort"):
    """BASELINE configs[4]: EER of n synthetic scores.  `value` = the north_star's algorithm (device radix sort + FAR/FRR sweep,
    or `method`), on the tie-free affine vector; the other method and two more input distributions ride along."""
    D = side_D
    out, first, cache = {}, None, {}
    for which in ("affine", "permuted", "sigmoid"):
        s, l = eer_inputs(ctx, n, which, cache)
        rec = {}
        for m in ("sort", "select"):
            ms, steps, clocks, res = ctx.timed_for(lambda m=m: D.eer_details(s, l, method=m), seconds / 3 if which != "affine" else seconds, min_steps=3)
            rec[m] = {"ms_per_step": ms / steps, "steps": steps, "scores_per_s": n * steps / (ms * 1e-3), "eer": res["eer"], "threshold": res["threshold"],
                      "eer_idx": res["eer_idx"], "clocks": clocks}
        rec["identical_result"] = all(rec["sort"][k] == rec["select"][k] for k in ("eer", "threshold", "eer_idx"))
        out[which] = rec
        if first is None:
            first = rec
        del s, l
        ctx.torch.cuda.empty_cache()
    main_, other = first[method], first["select" if method == "sort" else "sort"]
    gbs = 13.0 * main_["scores_per_s"] / 1e9
    roof = _hbm_roof(ctx, gbs, "13 B/score algorithmic (SURVEY.md 8d: read 4 B score + 1 B label, write 4 B sorted score + 4 B permutation); "
                               "implementation traffic: sort = 13 B prep + 16 B scatter per varying key byte (one kernel per pass: ticketed tiles, decoupled look-back) + 4 B sweep; "
                               "select = 5 B/score per varying key byte")
    return {"metric": "EER sweep (device radix sort + FAR/FRR crossing) on tie-free fp32 scores (BASELINE configs[4])",
            "value": main_["scores_per_s"], "unit": "scores/s", "ms_per_step": main_["ms_per_step"], "steps": main_["steps"], "n_scores": n,
            "dtype": "f32/f64", "eer_method": method, "roofline": roof, "eer": main_["eer"], "clocks": main_["clocks"],
            "eer_" + ("select" if method == "sort" else "sort"): {"value": other["scores_per_s"], "unit": "scores/s", "ms_per_step": other["ms_per_step"],
                                                                 "identical_result": first["identical_result"],
                                                                 "hbm_frac_at_13B_per_score": 13.0 * other["scores_per_s"] / 1e9 / ctx.pk["hbm_gbs"]},
            "inputs": {k: {"sort_ms": v["sort"]["ms_per_step"], "select_ms": v["select"]["ms_per_step"], "identical_result": v["identical_result"],
                           "eer": v["sort"]["eer"], "clocks_samples": v["sort"]["clocks"]["samples"]} for k, v in out.items()}}


def run_side_only(args, rank, world, local):
    """--workload cae | cnn1d | hybrid | eer: one side leg as a line of its own (profiling runs, A/B switches)."""
    import torch
    import torch.distributed as dist

    import dfs_b200 as D
    from dfs_b200 import synthetic as syn

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Ctx(args, torch, dist, rank, world, local)
    seconds = max(args.leg_seconds, 0.05)
    if args.workload == "eer":
        leg = leg_eer(ctx, D, seconds, args.eer_n, args.eer_method)
    else:
        P = min(args.pool, 16640) if args.workload != "cnn1d" else args.pool
        pool = D.fill_features(P, first_utt=rank * P, seed=1234, device=local)
        labels_global = torch.from_numpy(syn.labels(P * world)).to(ctx.dev)
        host_pool = None
        if args.workload == "hybrid":
            Pe = min(args.e2e_pool, P)
            host_pool = torch.empty((Pe, 321, 180), dtype=torch.float32, pin_memory=True)
            host_pool.copy_(pool[:Pe])
        side = Side(ctx, pool, labels_global, host_pool=host_pool)
        leg = {"cae": leg_cae, "cnn1d": leg_cnn1d, "hybrid": leg_hybrid}[args.workload](ctx, side, seconds)
    ctx.sampler.stop()
    if rank == 0:
        line = {"n_gpus": world, "warmup": 3, "higher_is_better": True, "scaling": "weak" if args.workload != "eer" else "replicas only",
                "vs_baseline": None, "data": "synthetic", "config": {"workload": args.workload, "l2": "inputs larger than L2"},
                "gpu_launches": int(D._native.launch_count())}
        line.update(leg)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # rank 0 prints ONE JSON line on stdout.  With NCCL_DEBUG=VERSION|WARN NCCL printf()s its version banner straight to
    # stdout: drop those two levels (errors still surface as exceptions); INFO / TRACE output goes to stderr
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        os.environ.pop("NCCL_DEBUG")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload != "cnn2d":
        run_side_only(args, rank, world, local)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import dfs_b200 as D
    from dfs_b200 import synthetic as syn
    from dfs_b200.distributed import gather_scores

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the scoring path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Ctx(args, torch, dist, rank, world, local)
    pk = ctx.pk

    P = args.pool
    sd = syn.cnn2d_state(0)
    scorer = D.Cnn2dScorer(sd, device=local, max_chunk=args.chunk)
    # rank r owns global utterances [r*P, (r+1)*P): generated on the device from (seed, global index)
    pool = D.fill_features(P, first_utt=rank * P, seed=1234, device=local)
    labels_global = torch.from_numpy(syn.labels(P * world)).to(dev)
    last = {}

    def step():
        s = scorer.score(pool, apply_sigmoid=True)
        g = gather_scores(s, n_total=P * world)      # one NCCL all-gather of 4 B/utterance over NVLink (no-op at N=1)
        last["s"], last["g"] = s, g
        return D.eer_details(g, labels_global)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        res = step()
    # kernel-time shares: one fully profiled step outside the timed region; inside it only the roofline kernel (conv3,
    # kernel id 2) carries event pairs, so the other launches run back to back
    scorer.set_option("profile", 1)
    step()
    share_ms, _ = scorer.profile(4)
    scorer.set_option("profile", 1 << 2)
    step()
    scorer.profile(4)
    kms, kcnt = [0.0] * 4, [0] * 4

    def collect():
        a, b = scorer.profile(4)
        for i in range(4):
            kms[i] += a[i]
            kcnt[i] += b[i]

    launches0 = D._native.launch_count()
    ms_max, clocks, res = ctx.timed(step, args.steps, warmup=0, after_step=collect)
    launches = D._native.launch_count() - launches0
    scorer.set_option("profile", 0)
    value = P * world * args.steps / (ms_max * 1e-3)
    s_last, g_last = last["s"], last["g"]

    # ---- sharded parity (N > 1): rank 0 re-scores the head of another rank's slice and compares with the gathered vector ----
    sharded = None
    if world > 1:
        ns = min(512, P)
        other = world - 1
        again = scorer.score(D.fill_features(ns, first_utt=other * P, seed=1234, device=local), apply_sigmoid=True)
        got = g_last[other * P:other * P + ns]
        sharded = {"sharded_equals_single": bool(torch.equal(again, got)), "max_abs_diff": float((again - got).abs().max()), "n": ns,
                   "note": f"rank 0 re-scored the first {ns} utterances of rank {other}'s slice on its own GPU and compared them with the "
                           "all-gathered vector (same kernels, another device, another position in the pass)"}

    # ---- e2e: the same metric through the public host-buffer call (pinned host -> H2D -> kernels -> D2H) ----
    Pe = min(args.e2e_pool, P)
    with D.hostmem.numa_local(local) as numa:                                       # pages next to this rank's GPU where the host has > 1 node
        host_pool = torch.empty((Pe, 321, 180), dtype=torch.float32, pin_memory=True)   # allocated after set_device, by this rank
        host_pool.copy_(pool[:Pe])
    lab_e = labels_global[:Pe]
    e2e_last = {}

    def e2e_step():
        e2e_last["s"] = scorer.score_host(host_pool, 1)
        return D.eer_details(e2e_last["s"], lab_e)

    e2e_ms, e2e_steps, e2e_clocks, _ = ctx.timed_for(e2e_step, args.e2e_seconds, min_steps=10, warmup=2)
    e2e_value = Pe * world * e2e_steps / (e2e_ms * 1e-3)
    e2e_scores = e2e_last["s"]
    ceiling_gbs = h2d_ceiling(ctx, host_pool)
    ceiling_utt = ceiling_gbs * 1e9 / BYTES_PER_UTT
    # the same call on an fp16 pinned slab (dfs_score_host_f16): the 2D-CNN's scores are bit-identical, the PCIe bytes halve
    host16 = host_pool.half().pin_memory()
    same16 = bool((scorer.score_host(host16, 1) == e2e_scores).all())
    e16_ms, e16_steps, _, _ = ctx.timed_for(lambda: D.eer_details(scorer.score_host(host16, 1), lab_e), args.e2e_seconds / 2, min_steps=5, warmup=1)
    e2e16_value = Pe * world * e16_steps / (e16_ms * 1e-3)
    del host16

    # ---- the other BASELINE configs as legs of the same line ----
    workloads = {}
    if not args.no_side:
        side = Side(ctx, pool, labels_global, c2=scorer, host_pool=host_pool)
        for name, fn in (("cae", leg_cae), ("hybrid", leg_hybrid), ("cnn1d", leg_cnn1d)):
            workloads[name] = fn(ctx, side, args.leg_seconds)
        del side
        if rank == 0:
            del pool                                             # make room for the 100 M-score vectors and their sort workspace
            torch.cuda.empty_cache()
            workloads["eer"] = leg_eer(ctx_single(ctx), D, args.leg_seconds, args.eer_n, args.eer_method)
            pool = None
        ctx.barrier()
    ctx.sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    chunk = args.chunk or 416
    conv3_ms = kms[2] / max(kcnt[2], 1)
    utt_per_launch = P / max(kcnt[2] / args.steps, 1)
    achieved = CONV3_FLOP_PER_UTT * utt_per_launch / (conv3_ms * 1e-3) / 1e12 if conv3_ms > 0 else 0.0
    traffic = None   # dram__bytes_read+write of that kernel from the committed ncu --set full capture, scaled to one launch
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)["dram_bytes_per_utterance"] * utt_per_launch
    whole = value / world * FLOP_PER_UTT["cnn2d"] / 1e12
    roofline = {"bound": "tensor", "kernel": "conv_tc_kernel<MODE_3X3S,64,128,N=256> (CNN2D conv3, 66% of the FLOPs)", "achieved": achieved,
                "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tflops_sustained"], "traffic": traffic,
                "frac_of_burst_peak": achieved / pk["tflops_burst"], "frac_of_nominal_peak": achieved / NOMINAL_TFLOPS,
                "peak_note": "`peak` is the cuBLAS bf16 GEMM measured back to back for seconds under the power cap (MEASURED_PEAKS.json): a reference "
                             "kernel, not a hardware ceiling, so a conv kernel that spends fewer joules per FLOP can read above 1.0 against it; "
                             "the burst (cold, best-of-10 cuBLAS) and nominal (datasheet 2,250) fractions are given beside it",
                "traffic_note": "DRAM bytes per launch (ncu); the tensor-bound kernel's algorithmic operand is the fp16 act2 read, 1.91 MB/utterance",
                "peak_source": pk["source"] + ", sustained (kernel timed inside a long step)",
                "flops_per_launch": CONV3_FLOP_PER_UTT * utt_per_launch, "avg_launch_ms": conv3_ms,
                "kernel_ms_share": {k: v / max(sum(share_ms), 1e-9) for k, v in zip(("xt_prep", "conv12_fused", "conv3", "head"), share_ms)},
                "kernel_ms_share_note": "from one fully profiled step before the timed region; conv3's launches are timed inside it; conv12_fused = blocks 1 + 2 "
                                        "of the 2D-CNN in one kernel (kernel ids 0..3 of dfs_model_profile)",
                "conv12_fused_tflops": (CONV2_FLOP_PER_UTT + 2 * 9 * 32 * 321 * 180) * P / (share_ms[1] * 1e-3) / 1e12 if share_ms[1] > 0 else None,
                "whole_path_tflops": whole, "whole_path_frac_of_sustained_peak": whole / pk["tflops_sustained"],
                "whole_path_frac_of_burst_peak": whole / pk["tflops_burst"], "whole_path_frac_of_nominal_peak": whole / NOMINAL_TFLOPS}

    out = {"metric": METRIC, "value": value, "unit": "utterances/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
           "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
           "data": "synthetic",
           "config": {"workload": WORKLOAD,
                      "utterances_per_step_per_gpu": P, "total_utterances": P * world * args.steps, "chunk": chunk,
                      "l2": "inputs larger than L2 (pool %.2f GB per GPU, cycled)" % (P * BYTES_PER_UTT / 1e9),
                      "weights": "random-init CNN2D, seeded (dfs_b200.synthetic.cnn2d_state(0)); no checkpoints ship with the reference",
                      "parallelism": f"dp{world} (utterance shards, one NCCL all-gather of scores per step)" if world > 1 else "dp1"},
           "eer": {"value": res["eer"], "threshold": res["threshold"], "n": P * world},
           "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline,
           "e2e": {"value": e2e_value, "unit": "utterances/s", "h2d_bytes_per_step": Pe * BYTES_PER_UTT, "d2h_bytes_per_step": Pe * 4,
                   "utterances_per_step_per_gpu": Pe, "steps": e2e_steps, "seconds": e2e_ms * 1e-3, "clocks": e2e_clocks,
                   "h2d_ceiling_gbs": ceiling_gbs, "h2d_ceiling_utterances_per_s": ceiling_utt, "frac_of_h2d_ceiling": e2e_value / ceiling_utt,
                   "frac_of_device_resident": e2e_value / value,
                   "h2d_ceiling_note": f"bare pinned cudaMemcpyAsync of the same pool in 96 MB pieces, {world} rank(s) copying at once, no kernels: "
                                       "what this host can feed; the fp32 e2e number is bound by it, not by a kernel",
                   "numa": dict(numa.applied, host_nodes=len(D.hostmem.host_nodes())),
                   "note": "dfs_score_host: pinned host features -> double-buffered H2D -> kernels -> D2H scores, + EER"},
           "e2e_f16_slab": {"value": e2e16_value, "unit": "utterances/s", "h2d_bytes_per_step": Pe * BYTES_PER_UTT // 2, "steps": e16_steps,
                            "scores_identical_to_fp32_slab": same16,
                            "note": "same call on an fp16 pinned slab (dfs_score_host_f16); informational: `e2e` above is the fp32 format "
                                    "the reference stores"},
           "workloads": workloads}
    if sharded is not None:
        out["parity"] = sharded

    if world == 1 and not args.no_cpu_baseline:
        n_cpu = 2048
        feats_dev = D.fill_features(n_cpu, first_utt=0, seed=1234, device=local) if pool is None else pool[:n_cpu]
        feats_cpu = feats_dev.cpu()
        rate, n_used, cores, ref_scores = cpu_reference_rate(torch, feats_cpu, sd, args.cpu_seconds)
        dev_scores = s_last[:n_used].cpu().numpy()
        rel = float(np.max(np.abs(dev_scores - ref_scores) / np.abs(ref_scores)))
        from oracle import eer as oeer
        # End-to-end EER check: labels correlated with the REFERENCE's score ranks (Bernoulli(sigmoid(6 (rank/n - 1/2))), seed 7)
        # so that the FAR/FRR crossing is sharp; with labels independent of the scores the curves are flat around the
        # crossing and the EER of 2,000 scores moves by 1e-3 under rank swaps far below the score tolerance.
        ranks = np.argsort(np.argsort(ref_scores, kind="stable"), kind="stable")
        prob = 1.0 / (1.0 + np.exp(-6.0 * (ranks / max(n_used, 1) - 0.5)))
        lab = (np.random.Generator(np.random.PCG64(7)).random(n_used) < prob).astype(np.uint8)
        eer_cpu, eer_gpu = oeer.calculate_eer(ref_scores, lab)[0], D.calculate_eer(dev_scores, lab)[0]
        out["cpu_baseline"] = {"value": rate, "unit": "utterances/s", "cores": cores, "kind": "port",
                               "sample": f"first {n_used} utterances of the pool, oracle port of the predict.py loop (bs 32, torch CPU fp32)"}
        out["parity"] = {"max_rel_err_scores_vs_cpu_reference": rel, "n": n_used, "tolerance": 1e-3,
                         "eer_cpu": eer_cpu, "eer_gpu": eer_gpu, "eer_delta_pp": 100.0 * abs(eer_cpu - eer_gpu),
                         "labels": "Bernoulli(sigmoid(6*(reference rank/n - 0.5))), seed 7",
                         "note": "random-init scores of i.i.d. utterances are ~4e-6 apart (sigmoid range 0.514-0.522): one rank swap moves the EER of "
                                 "2,000 scores by 0.05 pp; the 0.01 pp gate is held on the trained-like fixture (tests/test_gpu_round2.py, "
                                 "tests/golden/trained.npz), where logits span +-20"}
        # the same utterances through the full-fp32 CUDA-core kernels (Cnn2dScorer(precision="fp32"), csrc/cnn2d_fp32.cu): the
        # option for evaluations where the rank order of scores a few 1e-6 apart matters (random-init scores are)
        exact = D.Cnn2dScorer(sd, device=local, precision="fp32")
        exact.score(feats_dev[:16], apply_sigmoid=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s32 = exact.score(feats_dev[:n_used], apply_sigmoid=True).cpu().numpy()
        dt = time.perf_counter() - t0
        eer32 = D.calculate_eer(s32, lab)[0]
        out["parity"]["fp32_mode"] = {"max_rel_err_scores_vs_cpu_reference": float(np.max(np.abs(s32 - ref_scores) / np.abs(ref_scores))),
                                      "eer_gpu": eer32, "eer_delta_pp": 100.0 * abs(eer_cpu - eer32), "utterances_per_s": n_used / dt,
                                      "note": "precision=\"fp32\": fp32 operands and accumulation on the CUDA cores, explicit option"}
        del exact
        # ... and through the split-precision tensor-core kernels (Cnn2dScorer(precision="split")): every operand as fp16 value + residual,
        # three MMAs per product; rate over several passes of the same utterances, device-resident
        split = D.Cnn2dScorer(sd, device=local, precision="split")
        ssp = split.score(feats_dev[:n_used], apply_sigmoid=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            split.score(feats_dev[:n_used], apply_sigmoid=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 4
        ssp = ssp.cpu().numpy()
        eersp = D.calculate_eer(ssp, lab)[0]
        out["parity"]["split_mode"] = {"max_rel_err_scores_vs_cpu_reference": float(np.max(np.abs(ssp - ref_scores) / np.abs(ref_scores))),
                                       "eer_gpu": eersp, "eer_delta_pp": 100.0 * abs(eer_cpu - eersp), "utterances_per_s": n_used / dt,
                                       "note": "precision=\"split\": tcgen05 with fp16 value + residual operands (3 MMAs per product, fp32 accumulate), explicit option"}
        del split
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def ctx_single(ctx):
    """A view of the context for a leg that only rank 0 runs ("replicas only": the EER of one 100 M-score vector does not shard):
    no collectives in barrier / max."""
    class _One:
        pass
    one = _One()
    one.args, one.torch, one.dist, one.rank, one.world, one.local, one.dev, one.pk, one.sampler = (
        ctx.args, ctx.torch, ctx.dist, 0, 1, ctx.local, ctx.dev, ctx.pk, ctx.sampler)
    one.barrier = ctx.torch.cuda.synchronize
    one.max_over_ranks = lambda x: float(x)
    one.sum_over_ranks = lambda x: float(x)
    one.timed = lambda *a, **k: Ctx.timed(one, *a, **k)
    one.timed_for = lambda *a, **k: Ctx.timed_for(one, *a, **k)
    return one


if __name__ == "__main__":
    main()
