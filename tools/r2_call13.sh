#!/bin/bash
# round 2, GPU call 13: 3D with a column per cell (conflict-free shared access) vs tile bags vs the sorted kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_moments.py tests/test_host_replay.py tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2_call13_pytest.log 2>&1; tail -3 gpurun_out/r2_call13_pytest.log
EPB_BAG3D_CELLS=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "3 or thermal or relativ or mixed or reflect" > gpurun_out/r2_call13_pytest_bags.log 2>&1; tail -2 gpurun_out/r2_call13_pytest_bags.log
for cfg in "1 1 0" "1 1 1" "1 0 1" "0 0 0"; do
  set -- $cfg
  EPB_PUSH3D_VARIANT=$1 EPB_BAG3D_CELLS=$2 EPB_LOAD_MIXED=$3 timeout 600 python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_call13_c4_v$1_c$2_m$3.json 2> gpurun_out/r2_call13_c4_v$1_c$2_m$3.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2_call13_c4_v$1_c$2_m$3.json').read().strip().splitlines()[-1]); print('c4 share variant=$1 cells=$2 mixed=$3:', d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/r2_call13_c4_v$1_c$2_m$3.err').read()[-1200:])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:push_bag_3d -s 4 -c 1 -o gpurun_out/r2_prof_bag3d_cells -f \
  python bench.py --workload c4 --cells 192 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_call13_prof3d.log 2>&1
EPB_BENCH_E2E_BREAKDOWN=1 timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-parity-check > gpurun_out/r2_call13_c2.json 2> gpurun_out/r2_call13_c2.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_call13_c2.json').read().strip().splitlines()[-1]); print('c2:', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d.get('e2e_full'))"
grep "e2e breakdown" gpurun_out/r2_call13_c2.err
