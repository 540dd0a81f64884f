#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_call7_pytest.log 2>&1
tail -8 gpurun_out/r2_call7_pytest.log
timeout 600 python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_call7_c4_1gpu.json 2> gpurun_out/r2_call7_c4_1gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_call7_c4_1gpu.json').read().strip().splitlines()[-1]); print('c4 share 1 gpu:', d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
