"""profiles/rNN_hbm_ncu_full.md from an `ncu --set full` report of the HBM-class kernels (scripts/gpu_hbm_ncu.sh)."""
import csv, io, subprocess, sys

rep = sys.argv[1]
hbm_peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6468.0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if "average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]


def num(r, k, scale=1.0):
    try:
        return float(r[col[k]].replace(",", "")) * scale
    except Exception:
        return float("nan")


def to_bytes(r, k):
    u = units[col[k]]
    return num(r, k, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0))


def to_us(r, k):
    u = units[col[k]]
    return num(r, k, {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0))


print(f"# ncu --set full: the HBM-class kernels (`{rep.split('/')[-1]}`, per launch; cold-cache, serialised)\n")
print("Command: `bash scripts/gpu_hbm_ncu.sh` (second pass of `scripts/dev_hbm_once.py`: 16-window STFT / iSTFT batch of a 4-min stereo")
print("track, STFT-2048 feature passes, framewise RMS whole-track and all-chunks-in-one-launch, ZCR, YIN and LPC on 60 s).")
print(f"`achieved` = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration; `% of copy peak` against the measured {hbm_peak:.0f} GB/s.\n")
print("| # | kernel | grid x block | regs | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | % of copy peak | dram % (ncu) | L1/TEX % | L2 % | SM % | issue active % | warps active % | top stall reasons (warps per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for i, r in enumerate(rows[2:]):
    name = r[col["Kernel Name"]].replace("ac::", "").replace("void ", "").split("(")[0][:44]
    t = to_us(r, "gpu__time_duration.sum")
    rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / (t * 1e-6) / 1e9
    st = sorted(((num(r, h), h.split("issue_stalled_")[-1].replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)[:3]
    g = lambda k: (f"{num(r, k):.1f}" if k in col else "n/a")
    print(f"| {i} | {name} | {r[col['launch__grid_size']]} x {r[col['launch__block_size']]} | {r[col['launch__registers_per_thread']]} | {t:.1f} | "
          f"{rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | {100 * gbs / hbm_peak:.1f} | {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | "
          f"{g('l1tex__throughput.avg.pct_of_peak_sustained_elapsed')} | {g('lts__throughput.avg.pct_of_peak_sustained_elapsed')} | "
          f"{g('sm__throughput.avg.pct_of_peak_sustained_elapsed')} | {g('smsp__issue_active.avg.pct_of_peak_sustained_active')} | "
          f"{g('sm__warps_active.avg.pct_of_peak_sustained_active')} | " + ", ".join(f"{n} {v:.2f}" for v, n in st) + " |")
