#!/bin/bash
# round 2, GPU call 23: HC in the mover-buffer push; launch list of one bench run incl. the e2e loop (what the extra 2 ms per e2e step are)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hc_push.py -m gpu -q > gpurun_out/r2_call23_pytest.log 2>&1; tail -4 gpurun_out/r2_call23_pytest.log | cut -c1-250
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_call23_launches.csv \
  python bench.py --cells 2048 --steps 4 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e-full --no-mixed > gpurun_out/r2_call23_under_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2_call23_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i+2; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki][:50], float(r[vi].replace(',',''))/1e3) for r in rows[start:] if len(r)>vi]
print(len(seq),"launches")
# print the last ~60 launches (the e2e loop's last step)
for k,v in seq[-70:]: print(f"{k:52s} {v:10.1f} us")
PY
