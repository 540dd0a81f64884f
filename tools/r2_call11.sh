#!/bin/bash
# round 2, GPU call 11: tile-bag 3D kernel -- parity (all gpu tests), C4 per-GPU share with the new and the old kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_call11_pytest.log 2>&1
tail -30 gpurun_out/r2_call11_pytest.log | cut -c1-300
for v in 1 0; do
  EPB_PUSH3D_VARIANT=$v timeout 600 python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_call11_c4_v$v.json 2> gpurun_out/r2_call11_c4_v$v.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_call11_c4_v$v.json').read().strip().splitlines()[-1]); print('c4 share variant $v:', d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r2_call11_c4_v$v.err').read()[-1500:])"
done
EPB_LOAD_MIXED=1 timeout 600 python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_call11_c4_v1_mixed.json 2> gpurun_out/r2_call11_c4_v1_mixed.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_call11_c4_v1_mixed.json').read().strip().splitlines()[-1]); print('c4 share variant 1 mixed start:', d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
