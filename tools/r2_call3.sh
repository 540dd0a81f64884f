#!/bin/bash
# round 2, GPU call 3: where does the slot-column step spend its time?  per-kernel durations (ncu, serialised) of a
# 2048^2 bench + one full capture of push_slots_2d
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 160 --csv --log-file gpurun_out/r2_call3_launches.csv \
  python bench.py --cells 2048 --steps 4 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_call3_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:push_slots_2d -s 4 -c 1 -o gpurun_out/r2_prof_slots -f \
  python bench.py --cells 2048 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_call3_prof.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_deliver -s 4 -c 1 -o gpurun_out/r2_prof_deliver -f \
  python bench.py --cells 2048 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_call3_prof2.log 2>&1
ls -la gpurun_out/ | tail -5
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2_call3_launches.csv", errors="ignore")))
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
