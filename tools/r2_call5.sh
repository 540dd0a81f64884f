#!/bin/bash
# round 2, GPU call 5: kernel trims (deferred inbox reply, fewer FP64 / integer instructions) -- parity subset + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_parity.py tests/test_moments.py -m gpu -x -q > gpurun_out/r2_call5_pytest.log 2>&1
tail -6 gpurun_out/r2_call5_pytest.log
for cfg in "3 0" "3 1" "4 0"; do
  set -- $cfg
  EPB_SLOTS_MINB=$1 EPB_LOAD_MIXED=$2 timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-parity-check \
    > gpurun_out/r2_call5_bench_m$1_mix$2.json 2> gpurun_out/r2_call5_bench_m$1_mix$2.err
  echo "minb=$1 mixed=$2"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_call5_bench_m$1_mix$2.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2_call5_bench_m$1_mix$2.err").read()[-1500:])
PY
done
timeout 600 python bench.py --workload c3 --cells 1024 --steps 6 --warmup 3 --no-parity-check > gpurun_out/r2_call5_c3_1gpu.json 2> gpurun_out/r2_call5_c3_1gpu.err
tail -c 1500 gpurun_out/r2_call5_c3_1gpu.json; tail -5 gpurun_out/r2_call5_c3_1gpu.err
