"""Per-kernel counts of the SASS mnemonics that show which hardware paths a kernel uses (tcgen05 MMA / 2-CTA MMA / TMEM loads / TMA /
bulk copies / cluster barriers / 256-bit stores / packed FFMA2 / saturating fp16 packs).  Runs on the build host (no GPU):
    python tools/sass_evidence.py > profiles/<round>_sass_mnemonics.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR.2CTA.MULTICAST", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS.PHASECHK", "UCGABAR_ARV", "MAPA",
         "STG.E.ENL2.256", "FFMA2", "F2FP.SATFINITE", "HMMA", "DP4A", "MATCH", "VOTE", "SHFL", "REDUX", "ATOMS", "STS", "LDS"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def main():
    print("# cuobjdump -sass of the sm_100a objects: occurrences of selected mnemonics per kernel (static counts, not executions)")
    for obj in sorted(glob.glob(os.path.join(ROOT, "deep-fake-audio-classifier_b200", "build", "*.o"))):
        text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        fn, counts = None, collections.OrderedDict()
        for line in text.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                fn = m.group(1)
                counts[fn] = collections.Counter()
                continue
            if fn is None:
                continue
            m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            op = m.group(1)
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):
                    counts[fn][w] += 1
                    break
        shown = [(f, c) for f, c in counts.items() if any(c[w] for w in WATCH[:13])]
        if not shown:
            continue
        print(f"\n== {os.path.basename(obj)}")
        for f, c in shown:
            name = re.sub(r"\s+", " ", demangle(f))
            name = re.sub(r"\(.*", "", name)[:150]
            print(f"  {name}\n      " + "  ".join(f"{w}={c[w]}" for w in WATCH if c[w]))


if __name__ == "__main__":
    main()
