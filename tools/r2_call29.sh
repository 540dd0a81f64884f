#!/bin/bash
# round 2, GPU call 29: append API test, the bench as the driver runs it (e2e on a fresh load), reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_reference_binary_gpu.py -m gpu -q > gpurun_out/r2_call29_pytest.log 2>&1; tail -4 gpurun_out/r2_call29_pytest.log | cut -c1-250
( time timeout 900 python bench.py > gpurun_out/r2_call29_bench_default.json 2> gpurun_out/r2_call29_bench_default.err ) 2>&1 | grep real
python -c "
import json; d=json.loads(open('gpurun_out/r2_call29_bench_default.json').read().strip().splitlines()[-1]); print('default:', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['e2e']['value']/d['value'], 'kernel', d['roofline']['kernel_ms'], d['roofline']['frac'], 'traffic/alg', d['roofline']['traffic']/(d['roofline']['bytes_per_update']*d['config']['particles_total']), d['mixed_state']['ms_per_step'], d['parity_check'], d['gpu_launches'])"
tail -3 gpurun_out/r2_call29_bench_default.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_call29_bench_reference.json 2> gpurun_out/r2_call29_bench_reference.err ) 2>&1 | grep real
tail -c 900 gpurun_out/r2_call29_bench_reference.json
