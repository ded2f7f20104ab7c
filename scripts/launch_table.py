"""Print the per-launch durations of the LAST U-Net forward in an ncu `--metrics gpu__time_duration.sum --csv` log."""
import csv, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]; rows = rows[1:]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
first = [i for i, r in enumerate(rows) if "first_conv" in r[ki]]
seq = rows[first[-1]:]
tot = 0.0
for r in seq:
    name = r[ki].replace("ac::", "").replace("void ", "").split("(")[0]
    us = float(r[vi]) / 1e3
    tot += us
    print(f"{name[:44]:44s} {us:9.1f} us  grid {r[gi]}")
print(f"total {tot/1e3:.2f} ms over {len(seq)} launches")
