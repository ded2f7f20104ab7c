"""Turn an .ncu-rep (or a `--csv --log-file` launch list) into the small text summaries kept under profiles/.

  python scripts/ncu_summary.py rep  gpurun_out/prof.ncu-rep  > profiles/rNN_xxx_full.md
  python scripts/ncu_summary.py list gpurun_out/launches.csv  > profiles/rNN_xxx_launches.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__cycles_elapsed.max", "cycles"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma_inst_%"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_active_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def short(name):
    name = name.replace("ac::", "").replace("void ", "")
    return name.split("(")[0][:48]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    names = [k for k, _ in KEYS if k in col]
    print(f"# ncu --set full summary of `{path.split('/')[-1]}` (per launch; cold-cache, serialised)\n")
    print("| # | kernel | " + " | ".join(dict(KEYS)[k] + (f" [{units[col[k]]}]" if units[col[k]] else "") for k in names) + " |")
    print("|---|---|" + "---|" * len(names))
    for i, r in enumerate(rows[2:]):
        vals = []
        for k in names:
            v = r[col[k]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            vals.append(v)
        print(f"| {i} | {short(r[col['Kernel Name']])} | " + " | ".join(vals) + " |")


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        n = short(r[kn])
        t = float(r[mv].replace(",", ""))
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu launch list `{path.split('/')[-1]}`: gpu__time_duration.sum per kernel (ncu: cold-cache, serialised)\n")
    print("| kernel | launches | total ms | avg us | share |")
    print("|---|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {n} | {c} | {t / 1e6:.3f} | {t / c / 1e3:.1f} | {t / tot:.3f} |")
    print(f"\ntotal {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches")


if __name__ == "__main__":
    {"rep": rep, "list": launches}[sys.argv[1]](sys.argv[2])
