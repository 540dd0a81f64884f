#!/bin/bash
# round 2, GPU call 39: k_moment2_slots with div_rcp + row prefetch, div_rcp in the Nanbu-Perez pair operator:
# moment and collision tests, the config 5 bench line, and the launch times of the collision kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_moments.py tests/test_collisions.py -m gpu -q > gpurun_out/r2_call39_pytest.log 2>&1; tail -3 gpurun_out/r2_call39_pytest.log | cut -c1-300
timeout 400 python bench.py --workload c5 --no-e2e-full > gpurun_out/r2_call39_c5_1gpu.json 2> gpurun_out/r2_call39_c5_1gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_call39_c5_1gpu.json').read().strip().splitlines()[-1]); print('c5:', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'push', d['roofline']['kernel_ms'], d['collisions'], d['parity_check'])"
tail -3 gpurun_out/r2_call39_c5_1gpu.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_moment2|k_collide|k_settle|k_moment_post' -c 16 --csv --log-file gpurun_out/r2_call39_launches_collide.csv \
  python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e-full --no-mixed > gpurun_out/r2_call39_ncu.log 2>&1
grep -E "k_moment2|k_collide|k_settle" gpurun_out/r2_call39_launches_collide.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-120 | tail -16
