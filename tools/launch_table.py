"""Print the per-launch table of an ncu --metrics gpu__time_duration.sum CSV: python tools/launch_table.py file.csv [skip_fraction]"""
import csv
import sys
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
for r in rows[int(len(rows) * frac):]:
    print(f"{float(r['Metric Value'].replace(',', '')) / 1e3:9.1f} us  grid {r['Grid Size']:>14s}  {r['Kernel Name'][:70]}")
