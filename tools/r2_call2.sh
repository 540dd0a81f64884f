#!/bin/bash
# round 2, GPU call 2: slot-column layout -- parity first, then the bench against the sorted layout on the same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_call2_pytest.log 2>&1
tail -15 gpurun_out/r2_call2_pytest.log
for cfg in "5 3 0" "5 4 0" "3 3 0" "5 3 1" "3 3 1"; do
  set -- $cfg
  EPB_PUSH_VARIANT=$1 EPB_SLOTS_MINB=$2 EPB_LOAD_MIXED=$3 timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline $( [ "$cfg" = "5 3 0" ] || echo --no-parity-check ) \
    > gpurun_out/r2_call2_bench_v$1_m$2_mix$3.json 2> gpurun_out/r2_call2_bench_v$1_m$2_mix$3.err
  echo "variant=$1 minb=$2 mixed=$3"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_call2_bench_v$1_m$2_mix$3.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2_call2_bench_v$1_m$2_mix$3.err").read()[-1500:])
PY
done
