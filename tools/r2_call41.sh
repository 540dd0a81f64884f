#!/bin/bash
# round 2, GPU call 41: ncu --set full of the three slot-column moment passes and of k_collide after the div_rcp change (1024^2 x 64 ppc)
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_moment2_slots|k_collide' -c 4 -o gpurun_out/r2_prof_moments -f \
  python bench.py --workload c5 --cells 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e-full --no-mixed > gpurun_out/r2_call41.log 2>&1
ls -la gpurun_out/r2_prof_moments.ncu-rep; tail -1 gpurun_out/r2_call41.log | cut -c1-200
