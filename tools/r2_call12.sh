#!/bin/bash
# round 2, GPU call 12: A/B of the 2D arena layouts on one box, e2e breakdown, ncu of the 3D bag kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_collisions.py -m gpu -q -x > gpurun_out/r2_call12_pytest_rb1.log 2>&1; tail -2 gpurun_out/r2_call12_pytest_rb1.log
EPB_SLOTS_ROWBLOCK=0 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_collisions.py tests/test_moments.py -m gpu -q -x > gpurun_out/r2_call12_pytest_rb0.log 2>&1; tail -2 gpurun_out/r2_call12_pytest_rb0.log
for cfg in "1 0" "0 0" "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  EPB_SLOTS_ROWBLOCK=$1 EPB_LOAD_MIXED=$2 EPB_BENCH_E2E_BREAKDOWN=1 timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-parity-check \
    > gpurun_out/r2_call12_bench_rb$1_mix$2.json 2> gpurun_out/r2_call12_bench_rb$1_mix$2.err
  echo "rowblock=$1 mixed=$2"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_call12_bench_rb$1_mix$2.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
except Exception as e:
    print("failed", e)
PY
  grep "e2e breakdown" gpurun_out/r2_call12_bench_rb$1_mix$2.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:push_bag_3d -s 4 -c 1 -o gpurun_out/r2_prof_bag3d -f \
  python bench.py --workload c4 --cells 192 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r2_call12_prof3d.log 2>&1
ls -la gpurun_out/r2_prof_bag3d.ncu-rep
