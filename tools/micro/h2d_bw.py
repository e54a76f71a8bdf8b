"""Pinned-host -> device copy ceiling of this host with 1 ... N GPUs copying AT ONCE (one process per GPU).

    python tools/micro/h2d_bw.py                                              # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/micro/h2d_bw.py

Why: the end-to-end scoring number (bench.py `e2e`) is bound by how fast the host can feed fp32 features (231 KB per
utterance), and on the 8-GPU boxes of this pool that rate does not scale with the GPU count (round 1: 53 GB/s for one GPU,
112 GB/s for four, 184 GB/s for eight).  This prints, per allocation / affinity variant, the AGGREGATE GB/s of all ranks copying
concurrently -- the ceiling `e2e` is reported against -- so that what a builder can change (how the slab is allocated, which
cores the rank runs on, how many copies are in flight) is measured rather than guessed.

Variants: torch pin_memory | cudaHostAlloc portable (dfs_pinned_alloc) | + write-combined | rank pinned to its own share of the
cores before allocating and touching the slab (first-touch placement) | two copy streams per rank | 24 MB vs 96 MB vs 547 MB pieces.
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "deep-fake-audio-classifier_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dfs_b200 as D  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
UTT = 321 * 180 * 4
N_UTT = 4096                                   # 947 MB slab per rank
dst = torch.empty(2 * 2368 * UTT // 4, dtype=torch.float32, device=dev)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def measure(host, piece_utts, n_streams=1, seconds=0.6):
    """host: 1-D fp32 torch tensor over pinned memory.  Aggregate GB/s of all ranks, timed with CUDA events, max over ranks."""
    piece = piece_utts * UTT // 4
    n = host.numel()

    def sweep():
        k = 0
        for i in range(0, n, piece):
            m = min(piece, n - i)
            with torch.cuda.stream(streams[k % n_streams]):
                dst[(k & 1) * piece:(k & 1) * piece + m].copy_(host[i:i + m], non_blocking=True)
            k += 1

    sweep()
    barrier()
    t0 = time.perf_counter()
    sweep()
    torch.cuda.synchronize()
    reps = max(2, int(seconds / max(time.perf_counter() - t0, 1e-4)))
    t = torch.tensor([float(reps)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    reps = int(t.item())
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        sweep()
    for s in streams[:n_streams]:
        torch.cuda.current_stream().wait_stream(s)
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return n * 4 * reps * world / (float(ms.item()) * 1e-3) / 1e9


def torch_view(arr):
    return torch.from_numpy(arr.reshape(-1))


results = {}


def report(name, gbs):
    results[name] = gbs
    if rank == 0:
        print(f"{name:64s} {gbs:8.1f} GB/s aggregate  {gbs / world:7.1f} per GPU  = {gbs * 1e9 / UTT / 1e3:8.1f} k utt/s", flush=True)


if rank == 0:
    print(f"# pinned host -> device, {world} rank(s) copying at once; {os.cpu_count()} host cores; slab {N_UTT * UTT / 1e6:.0f} MB per rank", flush=True)
host_t = torch.empty(N_UTT * UTT // 4, dtype=torch.float32, pin_memory=True)
host_t.fill_(1.0)
for piece in (104, 416, 2368):
    report(f"torch pin_memory, pieces of {piece} utterances ({piece * UTT / 1e6:.0f} MB)", measure(host_t, piece))
report("torch pin_memory, 416-utterance pieces, 2 copy streams", measure(host_t, 416, n_streams=2))
del host_t
a = D.pinned_empty((N_UTT * UTT // 4,), "float32")
a[...] = 1.0
report("cudaHostAlloc portable (dfs_pinned_alloc), 416-utterance pieces", measure(torch_view(a), 416))
del a
wc = D.pinned_empty((N_UTT * UTT // 4,), "float32", write_combined=True)
wc[...] = 1.0
report("cudaHostAlloc write-combined, 416-utterance pieces", measure(torch_view(wc), 416))
del wc
# rank pinned to its own share of the host cores BEFORE the slab is allocated and first touched
cores = sorted(os.sched_getaffinity(0))
share = max(1, len(cores) // world)
mine = cores[rank * share:(rank + 1) * share] or cores
try:
    os.sched_setaffinity(0, mine)
    b = D.pinned_empty((N_UTT * UTT // 4,), "float32")
    b[...] = 1.0
    report(f"rank pinned to {len(mine)} core(s) before alloc + first touch, 416-utterance pieces", measure(torch_view(b), 416))
    del b
    os.sched_setaffinity(0, cores)
except OSError as e:
    if rank == 0:
        print("# sched_setaffinity not permitted:", e)

# ---- NUMA placement: where the pinned pages live relative to the GPU's PCIe root ---------------------------------------
from dfs_b200 import hostmem  # noqa: E402
from dfs_b200.hostmem import MPOL_BIND, MPOL_DEFAULT, MPOL_INTERLEAVE, host_nodes, node_cpus, set_mempolicy  # noqa: E402

node = hostmem.gpu_numa_node(local)
try:
    bdf = hostmem.gpu_pci_bdf(local)
except Exception as e:  # noqa: BLE001
    bdf = str(e)
nodes = host_nodes()
info = [None] * world
if world > 1:
    dist.all_gather_object(info, (rank, bdf, node))
else:
    info = [(rank, bdf, node)]
if rank == 0:
    print(f"# host NUMA nodes {nodes}; this process may run on cpus {sorted(os.sched_getaffinity(0))}")
    for n_ in nodes:
        print(f"#   node{n_} cpus {node_cpus(n_)[:4]}...({len(node_cpus(n_))})")
    for r_, b_, n_ in info:
        print(f"#   rank {r_} GPU {b_} numa_node {n_}")
for name, mode, sel in (("mempolicy BIND to the GPU's node", MPOL_BIND, [node] if node >= 0 else []),
                        ("mempolicy INTERLEAVE over all nodes", MPOL_INTERLEAVE, nodes)):
    if not sel or len(nodes) < 2:
        if rank == 0:
            print(f"# {name}: skipped (nodes {nodes}, gpu node {node})")
        continue
    err = set_mempolicy(mode, sel)
    errs = [None] * world
    if world > 1:
        dist.all_gather_object(errs, err)
    else:
        errs = [err]
    if any(errs):
        if rank == 0:
            print(f"# {name}: set_mempolicy failed, errno {errs}")
        set_mempolicy(MPOL_DEFAULT, [])
        continue
    local_cpus = [c for c in node_cpus(node) if c in cores] if mode == MPOL_BIND else []
    if local_cpus:
        os.sched_setaffinity(0, local_cpus)
    c_ = D.pinned_empty((N_UTT * UTT // 4,), "float32")
    c_[...] = 1.0
    report(f"{name} (+ cpu affinity to {len(local_cpus)} local cores), cudaHostAlloc, 416-utterance pieces", measure(torch_view(c_), 416))
    del c_
    set_mempolicy(MPOL_DEFAULT, [])
    os.sched_setaffinity(0, cores)

with hostmem.numa_local(local) as nl:
    e_ = D.pinned_empty((N_UTT * UTT // 4,), "float32")
    e_[...] = 1.0
applied = [None] * world
if world > 1:
    dist.all_gather_object(applied, nl.applied)
else:
    applied = [nl.applied]
report("hostmem.numa_local(device) around pinned_empty + first touch (the product call)", measure(torch_view(e_), 416))
if rank == 0:
    print("#   applied per rank:", applied)
del e_

# ---- how the aggregate builds up: only the first k ranks copy, the others idle (same slab, default placement) ----------
if world > 1:
    d_ = D.pinned_empty((N_UTT * UTT // 4,), "float32")
    d_[...] = 1.0
    dv = torch_view(d_)
    k = 1
    while k <= world:
        active = rank < k
        piece = 416 * UTT // 4
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        reps = 12 if active else 0
        for _ in range(reps):
            for kk, i in enumerate(range(0, dv.numel(), piece)):
                m = min(piece, dv.numel() - i)
                dst[(kk & 1) * piece:(kk & 1) * piece + m].copy_(dv[i:i + m], non_blocking=True)
        ev1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([ev0.elapsed_time(ev1) if active else 0.0], dtype=torch.float64, device=dev)
        per = [torch.zeros_like(ms) for _ in range(world)]
        dist.all_gather(per, ms)
        rates = [dv.numel() * 4 * 12 / (float(p.item()) * 1e-3) / 1e9 for p in per[:k]]
        results[f"first {k} rank(s) copying"] = rates
        if rank == 0:
            print(f"first {k} rank(s) copying, others idle: per-rank GB/s {[round(r_, 1) for r_ in rates]}  sum {sum(rates):.1f}", flush=True)
        k *= 2
    del d_, dv
if rank == 0:
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"h2d_bw_n{world}.json"), "w") as f:
        json.dump({"world": world, "cores": os.cpu_count(), "gbs_aggregate": results}, f, indent=1)
if world > 1:
    dist.destroy_process_group()
