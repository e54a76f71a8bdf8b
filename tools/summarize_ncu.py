"""Turn the ncu artefacts brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python tools/summarize_ncu.py <tag>      # reads gpurun_out/launches.csv and gpurun_out/prof_conv.ncu-rep
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active"]


KEYS += ["dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
         "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
         "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "lts__t_sector_hit_rate.pct"]

LISTS = (("launches.csv", "launch_share", "bench.py (2D-CNN headline workload)"),
         ("launches_hybrid.csv", "hybrid_launch_share", "bench.py --workload hybrid (2D-CNN + 1D-CNN + CAE-MSE + blend + EER)"),
         ("launches_cae.csv", "cae_launch_share", "bench.py --workload cae"),
         ("eer_launches.csv", "eer_launch_share", "bench.py --workload eer (100 M scores: sort path, then select path)"))
REPORTS = (("prof_conv.ncu-rep", "conv_kernels_ncu_full", "2D-CNN conv kernels"),
           ("prof_hybrid.ncu-rep", "hybrid_kernels_ncu_full", "CAE enc1 / final and 1D-CNN fused layer 1"),
           ("prof_eer.ncu-rep", "eer_kernels_ncu_full", "EER sort path: radix count and scatter passes"),
           ("prof_sel.ncu-rep", "eer_select_kernels_ncu_full", "EER select path: TMA-fed (digit, label) histogram, first and later levels"),
           ("prof_c1d.ncu-rep", "cnn1d_fused_and_prep_ncu_full", "1D-CNN fused layer 1, transposing input prep, CAE score finish"),
           ("prof_cae.ncu-rep", "cae_layers_ncu_full", "CAE layers of one pass (prep, enc1, enc2 PAIR, enc3, enc4 as 4 groups of N = 64, dec1 / dec2 wide, dec3 + final + MSE)"))


def launch_share(tag):
    for fname, suffix, what in LISTS:
        path = os.path.join(GP, fname)
        if not os.path.exists(path):
            continue
        rows = [r for r in csv.reader(open(path)) if len(r) > 5]
        hdr = rows[0]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        tot, cnt = collections.defaultdict(float), collections.Counter()
        for r in rows[1:]:
            try:
                v = float(r[vi].replace(",", ""))
            except ValueError:
                continue
            v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
            name = re.sub(r"\(.*", "", r[ki])[:100]
            tot[name] += v
            cnt[name] += 1
        s = sum(tot.values())
        with open(os.path.join(OUT, f"{tag}_{suffix}.txt"), "w") as f:
            f.write(f"# {what}\n")
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
            f.write(f"# total {s:.1f} us over {sum(cnt.values())} profiled launches\n")
            for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
                f.write(f"{v / s * 100:6.2f}%  {v:10.1f} us  n={cnt[k]:4d}  avg {v / cnt[k]:8.1f} us  {k}\n")


def full_set(tag):
    for fname, suffix, what in REPORTS:
        rep = os.path.join(GP, fname)
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        with open(os.path.join(OUT, f"{tag}_{suffix}.txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on: {what} (one launch each)\n")
            for r in rows[2:]:
                f.write(f"\n== {r[hdr.index('Kernel Name')]}\n")
                for i, h in enumerate(hdr):
                    if any(h.endswith(k) for k in KEYS) and r[i] not in ("",):
                        f.write(f"  {h:95s} {units[i]:16s} {r[i]}\n")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launch_share(tag)
    full_set(tag)
    print(sorted(os.listdir(OUT)))
