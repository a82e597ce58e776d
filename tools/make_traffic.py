"""profiles/iterate_traffic.json from the ncu captures of the iteration kernel (profiles/<tag>_ncu_iterate_*.txt, written by
tools/make_profiles.py): dram__bytes_read + dram__bytes_write of ONE launch, keyed by the digest of the kernel's sources so
that bench.py only quotes a figure captured from the build it is running.    python tools/make_traffic.py r02b"""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (source_digest)

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram_bytes(path):
    tot = 0.0
    for ln in open(path):
        m = re.match(r"\s*dram__bytes_(read|write)\.sum\s+([\d.,]+)\s+(\w+)", ln)
        if m:
            tot += float(m.group(2).replace(",", "")) * UNIT[m.group(3)]
    return int(tot)


def main(tag):
    files = {"f32_dz500_B128": f"{tag}_ncu_iterate_f32.txt", "f64_dz500_B128": f"{tag}_ncu_iterate_f64.txt",
             "f32_dz1000_B128": f"{tag}_ncu_iterate_f32_dz1000.txt"}
    entries = {}
    for key, f in files.items():
        p = os.path.join(ROOT, "profiles", f)
        if os.path.exists(p):
            entries[key] = dram_bytes(p)
    out = {"digest": bench.source_digest(), "source": f"ncu --set full captures profiles/{tag}_ncu_iterate_*.txt (one iterate_kernel launch each: "
           "61 ADMM passes, 4 stop checks at dz=500, 3 at dz=1000)", "entries": entries}
    with open(os.path.join(ROOT, "profiles", "iterate_traffic.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
