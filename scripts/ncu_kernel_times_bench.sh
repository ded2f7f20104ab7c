# usage: bash scripts/ncu_kernel_times_bench.sh <kernel-regex> [count]  -> per-launch gpu__time_duration of matching kernels in one bench step
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$1 -c ${2:-8} --csv --log-file /tmp/kt.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(l for l in open('/tmp/kt.csv') if l.startswith('"')))
h = rows[0]; k, v = h.index("Kernel Name"), h.index("Metric Value")
for r in rows[1:]:
    print(f"{r[k].replace('ac::','').replace('void ','').split('(')[0][:40]:40s} {float(r[v])/1e3:8.1f} us")
PY
