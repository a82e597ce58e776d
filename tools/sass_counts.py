"""Here (no GPU): per-kernel counts of the SASS mnemonics that prove the Blackwell paths of the shipped library
(tcgen05.mma / commit / ld, TMEM allocation, bulk TMA, mbarriers, fp64 DMMA, packed FFMA2) -> profiles/<tag>_sass_counts.txt.
Usage: python tools/sass_counts.py <tag>"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
so = os.path.join(ROOT, "lqp_py_b200", "_lqpb.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
mn = ["UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UBLKCP", "UBLKPF", "SYNCS", "DMMA", "FFMA2"]
cur, counts = None, collections.OrderedDict()
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
    elif cur:
        for k in mn:
            if k in line:
                counts[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.split("\n")
out = ["# cuobjdump -sass lqp_py_b200/_lqpb.so (sm_100a): instruction counts per kernel of the mnemonics that prove the Blackwell paths",
       "#   UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld (TMEM -> registers), UTCATOMSWS = tcgen05.alloc / dealloc,",
       "#   UBLKCP = cp.async.bulk (1-D TMA), UBLKPF = cp.async.bulk.prefetch.L2, SYNCS = mbarrier operations, DMMA = fp64 mma.sync,",
       "#   FFMA2 = packed fp32 FMA",
       "# library stamp (sha256 of the sources + flags): " + open(os.path.join(ROOT, "lqp_py_b200", "_lqpb.stamp")).read().strip(), ""]
tot = collections.Counter()
for (k, c), nm in zip(counts.items(), names):
    if sum(c.values()):
        tot.update(c)
        out.append(f"{nm[:100]:100s} " + "  ".join(f"{m}={c[m]}" for m in mn if c[m]))
out += ["", "TOTAL  " + "  ".join(f"{m}={tot[m]}" for m in mn)]
path = os.path.join(ROOT, "profiles", f"{tag}_sass_counts.txt")
open(path, "w").write("\n".join(out) + "\n")
print(path, "\n", out[-1])
