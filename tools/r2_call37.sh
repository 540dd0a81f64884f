#!/bin/bash
# round 2, GPU call 37: launch list of the config 5 step (what epb_collide's time is made of)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_call37_launches_c5.csv \
  python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e-full --no-mixed > gpurun_out/r2_call37.log 2>&1
tail -2 gpurun_out/r2_call37.log | cut -c1-300
wc -l gpurun_out/r2_call37_launches_c5.csv
