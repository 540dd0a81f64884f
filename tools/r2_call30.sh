#!/bin/bash
# round 2, GPU call 30: config 5 (collisions every step) bench line, collision tests after the tolerance change
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_collisions.py -m gpu -q > gpurun_out/r2_call30_pytest.log 2>&1; tail -2 gpurun_out/r2_call30_pytest.log | cut -c1-200
timeout 900 python bench.py --workload c5 --no-e2e-full > gpurun_out/r2_call30_c5_1gpu.json 2> gpurun_out/r2_call30_c5_1gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_call30_c5_1gpu.json').read().strip().splitlines()[-1]); print('c5:', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'push', d['roofline']['kernel_ms'], d['collisions'], d['parity_check'])"
tail -3 gpurun_out/r2_call30_c5_1gpu.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 2 -c 1 -o gpurun_out/r2_prof_collide -f \
  python bench.py --workload c5 --cells 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e-full --no-mixed > gpurun_out/r2_call30_prof.log 2>&1
ls -la gpurun_out/r2_prof_collide.ncu-rep
