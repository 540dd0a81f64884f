#!/bin/bash
# round 2, GPU call 28: why the 64-thread kernel on 16x4 tiles is slower (ncu --set full, 2048^2)
mkdir -p gpurun_out
EPB_DEBUG=1 EPB_SLOTS_CTY=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:push_slots_2d -s 4 -c 1 -o gpurun_out/r2_prof_slots_cty4 -f \
  python bench.py --cells 2048 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e-full --no-mixed > gpurun_out/r2_call28_prof.log 2>&1
ls -la gpurun_out/r2_prof_slots_cty4.ncu-rep
