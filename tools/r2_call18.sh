#!/bin/bash
# round 2, GPU call 18: 3D slot columns with warp-private deposit tiles (16x4x3 tiles, 6 warps, no shared atomics on the main path)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_moments.py tests/test_thermal_bc.py tests/test_gpu_bench_parity.py -m gpu -q -x -k "3 or 3d or three" > gpurun_out/r2_call18_pytest.log 2>&1; tail -3 gpurun_out/r2_call18_pytest.log
for cfg in "1 0" "1 1"; do
  set -- $cfg
  EPB_PUSH3D_VARIANT=$1 EPB_LOAD_MIXED=$2 timeout 600 python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_call18_c4_v$1_m$2.json 2> gpurun_out/r2_call18_c4_v$1_m$2.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_call18_c4_v$1_m$2.json').read().strip().splitlines()[-1]); print('c4 share variant=$1 mixed=$2:', d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d.get('parity_check'))
except Exception as e:
    print('failed', e); print(open('gpurun_out/r2_call18_c4_v$1_m$2.err').read()[-1200:])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:push_bag_3d -s 4 -c 1 -o gpurun_out/r2_prof_bag3d_v3 -f \
  python bench.py --workload c4 --cells 192 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_call18_prof3d.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
