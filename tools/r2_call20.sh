#!/bin/bash
# round 2, GPU call 20: full GPU suite on the current library; 2D bench A/B on one box: deferred inbox write (new) vs the previous library
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_call20_pytest.log 2>&1; tail -5 gpurun_out/r2_call20_pytest.log | cut -c1-250
cp epoch_b200/libepoch_b200.so /tmp/new.so
run() {  # name, mixed
  EPB_LOAD_MIXED=$2 timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-parity-check --no-e2e-full \
    > gpurun_out/r2_call20_$1.json 2> gpurun_out/r2_call20_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_call20_$1.json").read().strip().splitlines()[-1])
    print("$1", d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/r2_call20_$1.err").read()[-1500:])
PY
}
run new_mix0 0
cp build_ab/libepoch_b200_old.so epoch_b200/libepoch_b200.so
run old_mix0 0
run old_mix1 1
cp /tmp/new.so epoch_b200/libepoch_b200.so
run new_mix1 1
run new_mix0_again 0
