ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"viterbi|yin_probs|backtrack|pyin_bp" -c 10 --csv --log-file /tmp/kt.csv python scripts/dev_pyin_paths.py /tmp/x.npz > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(l for l in open('/tmp/kt.csv') if l.startswith('"')))
h = rows[0]; k, v = h.index("Kernel Name"), h.index("Metric Value")
for r in rows[1:]:
    print(f"{r[k].replace('ac::','').replace('void ','').split('(')[0][:44]:44s} {float(r[v])/1e6:8.2f} ms")
PY
