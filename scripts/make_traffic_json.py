"""profiles/rNN_unet_ncu_full.md (scripts/ncu_summary.py rep) -> profiles/rNN_traffic.json: measured DRAM bytes per launch of the
dominant kernel class (`roofline.traffic` in the bench line = dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full).

    python scripts/make_traffic_json.py profiles/r02_unet_ncu_full.md profiles/r02_traffic.json
"""
import json
import sys
from collections import OrderedDict

src, dst = sys.argv[1], sys.argv[2]
rows = [l for l in open(src) if l.startswith("| ") and not l.startswith("| #") and not l.startswith("|---")]
hdr = [h.strip() for h in [l for l in open(src) if l.startswith("| #")][0].strip().strip("|").split("|")]
i_k = hdr.index("kernel")
i_rd = [i for i, h in enumerate(hdr) if h.startswith("dram_rd")][0]
i_wr = [i for i, h in enumerate(hdr) if h.startswith("dram_wr")][0]
unit_rd = hdr[i_rd].split("[")[1].rstrip("]")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
per = OrderedDict()
total, n = 0.0, 0
for l in rows:
    c = [x.strip() for x in l.strip().strip("|").split("|")]
    if "tc_conv3x3" not in c[i_k]:
        continue
    b = (float(c[i_rd]) + float(c[i_wr])) * scale[unit_rd]
    per.setdefault(c[i_k], []).append(round(b / 1e6, 1))
    total += b
    n += 1
out = {"conv3x3_tcgen05": {"dram_bytes_per_launch": total / max(n, 1), "launches_per_forward": n, "batch_windows": 16,
                           "per_kernel_MB": per,
                           "source": f"ncu --set full --clock-control none, {src} (all conv launches of one 16-window forward: "
                                     "dram__bytes_read.sum + dram__bytes_write.sum)"}}
json.dump(out, open(dst, "w"), indent=1)
print(dst, n, "launches,", round(total / max(n, 1) / 1e6, 1), "MB per launch")
