"""Turn the scratch files a GPU session left in gpurun_out/ into the tracked evidence under profiles/.

    python tools/make_profiles.py r01a          # tag of the round / session

Copies the ncu launch list(s) and bench JSON lines, writes one text summary per .ncu-rep (the metrics
of tools/ncu_summary.py) and a per-kernel share table of every launch list.  Reads .ncu-rep files with
`ncu -i` (works without a GPU).
"""
import collections
import csv
import glob
import io
import os
import shutil
import subprocess
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402


def shares(path):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(r["Kernel Name"].split("(")[0][:80], [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", ""))
    tot = sum(v[1] for v in agg.values()) or 1.0
    out = [f"# {os.path.basename(path)}: {len(rows)} launches, gpu__time_duration.sum total {tot / 1e3:.1f} us "
           f"(cold-cache, serialised under ncu: compare SHARES, not absolutes)"]
    for k, v in sorted(agg.items(), key=lambda t: -t[1][1]):
        out.append(f"{v[0]:5d} launches {v[1] / 1e3:10.1f} us {100 * v[1] / tot:5.1f}%  {k}")
    return "\n".join(out) + "\n"


def main(tag):
    src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    for f in sorted(glob.glob(os.path.join(src, "launches_*.csv"))):
        base = os.path.basename(f)
        shutil.copy(f, os.path.join(dst, f"{tag}_{base}"))
        with open(os.path.join(dst, f"{tag}_{base[:-4]}_shares.txt"), "w") as fh:
            fh.write(shares(f))
    for f in sorted(glob.glob(os.path.join(src, "bench_*.json"))):
        if os.path.getsize(f):
            shutil.copy(f, os.path.join(dst, f"{tag}_{os.path.basename(f)}"))
    for f in sorted(glob.glob(os.path.join(src, "*.ncu-rep"))):
        buf = io.StringIO()
        with redirect_stdout(buf):
            ncu_summary.main(f)
        name = os.path.basename(f)[:-8]
        with open(os.path.join(dst, f"{tag}_ncu_{name}.txt"), "w") as fh:
            fh.write(f"# ncu --set full --clock-control none --import-source on, {name}.ncu-rep "
                     f"(numbers under the profiler: not bench values)\n")
            fh.write(buf.getvalue())
    print("\n".join(sorted(os.listdir(dst))))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
