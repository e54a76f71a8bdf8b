"""Where the pinned feature table lives relative to the GPU that reads it.

The end-to-end scoring rate (``dfs_score_host`` / ``dfs_group_score_host``) is bound by the pinned-host -> device copy of
231 KB per utterance (the reference re-reads its pickled table per model, src/predict.py:100-111; here the table is one
pinned slab, dropin/ingest.py).  On a two-socket host with several GPUs a slab whose pages sit on the other socket crosses
the inter-socket link on every copy, so the aggregate rate stops scaling with the GPU count.  ``numa_local(device)`` makes
the calling thread allocate (and run) next to that GPU's PCIe root while a slab is allocated and first touched:

    with hostmem.numa_local(torch.cuda.current_device()):
        slab = dfs_b200.pinned_empty((n, 321, 180))

It is a no-op (and says so in ``.applied``) on single-node hosts and where the container forbids ``set_mempolicy``.
Linux x86-64 only (syscall number); no libnuma dependency.
"""
from __future__ import annotations

import ctypes as C
import os
import re

MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND, MPOL_INTERLEAVE = 0, 1, 2, 3
_NR_SET_MEMPOLICY = 238                      # x86-64
_libc = None


def _syscall():
    global _libc
    if _libc is None:
        _libc = C.CDLL(None, use_errno=True)
    return _libc.syscall


def host_nodes():
    """NUMA node ids of this host ([] if /sys is not readable)."""
    try:
        return sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if re.fullmatch(r"node\d+", d))
    except OSError:
        return []


def parse_cpulist(txt):
    """"0-3,8,10-11" -> [0, 1, 2, 3, 8, 10, 11]"""
    out = []
    for part in txt.strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def node_cpus(node):
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            return parse_cpulist(f.read())
    except OSError:
        return []


def gpu_pci_bdf(device):
    import torch
    pr = torch.cuda.get_device_properties(device)
    return f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"


def gpu_numa_node(device):
    """NUMA node of the GPU's PCIe root from sysfs, or -1 (unknown / single node / virtualised)."""
    try:
        with open(f"/sys/bus/pci/devices/{gpu_pci_bdf(device)}/numa_node") as f:
            return int(f.read().strip())
    except (OSError, ValueError, RuntimeError, AttributeError):
        return -1


def set_mempolicy(mode, nodes=()):
    """Thread memory policy; returns 0 or the errno (EPERM under a seccomp profile that filters the call)."""
    mask = C.c_ulong(0)
    for n in nodes:
        mask.value |= 1 << n
    r = _syscall()(_NR_SET_MEMPOLICY, mode, C.byref(mask) if nodes else None, 64 if nodes else 0)
    return 0 if r == 0 else C.get_errno()


class numa_local:
    """Context manager: memory policy PREFERRED = the GPU's node + CPU affinity to that node's cores (of those this process may
    use), restored on exit.  ``applied`` tells what took effect: {"node", "mempolicy", "cpus"}."""

    def __init__(self, device, policy="bind"):          # "bind" (prefer the GPU's node, run on its cores) | "interleave"
        self.device, self.policy = device, policy
        self.applied = {"node": -1, "mempolicy": False, "cpus": 0}
        self._cores = None

    def __enter__(self):
        nodes = host_nodes()
        node = gpu_numa_node(self.device)
        self.applied["node"] = node
        if len(nodes) < 2 or (node < 0 and self.policy == "bind"):
            return self
        if self.policy == "interleave":
            ok = set_mempolicy(MPOL_INTERLEAVE, nodes) == 0
        else:
            ok = set_mempolicy(MPOL_PREFERRED, [node]) == 0      # PREFERRED, not BIND: falls back to other nodes instead of failing when the node is full
        self.applied["mempolicy"] = ok
        if self.policy == "bind":
            try:
                cores = sorted(os.sched_getaffinity(0))
                local = [c for c in node_cpus(node) if c in cores]
                if local:
                    os.sched_setaffinity(0, local)
                    self._cores = cores
                    self.applied["cpus"] = len(local)
            except OSError:
                pass
        return self

    def __exit__(self, *exc):
        if self.applied["mempolicy"]:
            set_mempolicy(MPOL_DEFAULT)
        if self._cores is not None:
            try:
                os.sched_setaffinity(0, self._cores)
            except OSError:
                pass
        return False
