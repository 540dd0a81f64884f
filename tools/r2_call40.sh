#!/bin/bash
# round 2, GPU call 40: the whole GPU suite, smoke() and the default bench line on the final library
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r2_call40_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_call40_pytest_gpu.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/r2_call40_bench_c2_1gpu.json 2> gpurun_out/r2_call40_bench_c2_1gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_call40_bench_c2_1gpu.json').read().strip().splitlines()[-1]); print('c2:', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], d['parity_check'], d.get('clocks'))"
tail -2 gpurun_out/r2_call40_bench_c2_1gpu.err
