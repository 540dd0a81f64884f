#!/bin/bash
# round 2, GPU call 21: pipelined e2e (async scalars, non-blocking source planes), the bench exactly as the driver runs it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_reference_binary_gpu.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_call21_pytest.log 2>&1; tail -4 gpurun_out/r2_call21_pytest.log | cut -c1-250
( time timeout 900 python bench.py > gpurun_out/r2_call21_bench_default.json 2> gpurun_out/r2_call21_bench_default.err ) 2>&1 | grep real
tail -c 3500 gpurun_out/r2_call21_bench_default.json; tail -3 gpurun_out/r2_call21_bench_default.err
EPB_BENCH_E2E_BREAKDOWN=1 timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-parity-check --no-e2e-full --no-mixed > gpurun_out/r2_call21_bench_bd.json 2> gpurun_out/r2_call21_bench_bd.err
grep "e2e breakdown" gpurun_out/r2_call21_bench_bd.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_call21_bench_bd.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['value']/d['value'])"
