"""Top warp-stall sampling sites of one kernel in an ncu report (SASS level, with the stall-reason columns that dominate).

    python tools/ncu_hotspots.py <report.ncu-rep> <substring of the kernel name> [min share %]"""
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    share = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif cur is not None and r:
            cur["rows"].append(r)
    seen = set()
    for k in kernels:
        if pat not in k["name"] or k["name"] in seen:
            continue
        seen.add(k["name"])
        hdr, body = k["rows"][0], k["rows"][1:]
        si, src = hdr.index("# Samples"), hdr.index("Source")
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or h.startswith("Stall")]
        tot = sum(int(r[si]) for r in body if r[si].isdigit())
        print(k["name"][:140], "| samples", tot)
        for idx, r in enumerate(body):
            n = int(r[si]) if r[si].isdigit() else 0
            if n > tot * share / 100:
                why = sorted(((int(r[i]), hdr[i]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:2]
                print(f"{idx:5d} {n:6d} {100 * n / tot:5.1f}%  {r[src].strip()[:90]:90s} {why}")


if __name__ == "__main__":
    main()
