#!/bin/bash
# round 2, GPU call 4: group inboxes -- parity, then bench with / without them, fresh and mixed start
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_call4_pytest.log 2>&1
tail -12 gpurun_out/r2_call4_pytest.log
for cfg in "1 0" "0 0" "1 1"; do
  set -- $cfg
  EPB_SLOTS_INBOX=$1 EPB_LOAD_MIXED=$2 timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline $( [ "$cfg" = "1 0" ] || echo --no-parity-check ) \
    > gpurun_out/r2_call4_bench_inbox$1_mix$2.json 2> gpurun_out/r2_call4_bench_inbox$1_mix$2.err
  echo "inbox=$1 mixed=$2"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_call4_bench_inbox$1_mix$2.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d.get("parity_check"), d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2_call4_bench_inbox$1_mix$2.err").read()[-1500:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 160 --csv --log-file gpurun_out/r2_call4_launches.csv \
  python bench.py --cells 2048 --steps 4 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_call4_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:push_slots_2d -s 4 -c 1 -o gpurun_out/r2_prof_slots_inbox -f \
  python bench.py --cells 2048 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_call4_prof.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2_call4_launches.csv", errors="ignore")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
if hdr:
    h = rows[hdr[0]]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr[0] + 2:]:
        if len(r) > vi:
            k = r[ki][:60]; agg.setdefault(k, []).append(float(r[vi].replace(",", "")))
    for k, v in agg.items():
        print(f"{k:60s} n={len(v):4d} total_ms={sum(v)/1e6:9.3f} mean_us={sum(v)/len(v)/1e3:9.1f}")
PY
