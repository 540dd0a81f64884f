#!/bin/bash
# round 2, GPU call 36: slot-column moment kernel (k_moment2_slots) -- moment and collision parity tests, then the config 5 bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_moments.py tests/test_collisions.py -m gpu -q > gpurun_out/r2_call36_pytest.log 2>&1; tail -3 gpurun_out/r2_call36_pytest.log | cut -c1-300
timeout 400 python bench.py --workload c5 --no-e2e-full > gpurun_out/r2_call36_c5_1gpu.json 2> gpurun_out/r2_call36_c5_1gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_call36_c5_1gpu.json').read().strip().splitlines()[-1]); print('c5:', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'push', d['roofline']['kernel_ms'], d['collisions'], d['parity_check'])"
tail -3 gpurun_out/r2_call36_c5_1gpu.err
