#!/bin/bash
# round 2, GPU call 34: what the driver runs at round end, on the final tree: build check, GPU suite, smoke, bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); print('build ok')" 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_call34_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2_call34_pytest_gpu.txt | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_call34_bench.json 2> gpurun_out/r2_call34_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_call34_bench.json').read().strip().splitlines()[-1]); print('bench:', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['roofline']['frac'], d['parity_check'], d['gpu_launches'], sorted(d.keys()))"
