#!/bin/bash
# round 2, GPU call 27: cp.async row prefetch (EPB_SLOTS_PREFETCH=1) and the 64-thread / 144-register kernel on 16x4 tiles
# (EPB_SLOTS_CTY=4): parity, then A/B against the default on one box; e2e loop with warm-up iterations
mkdir -p gpurun_out
for cfg in "EPB_SLOTS_PREFETCH=1" "EPB_SLOTS_CTY=4"; do
  env EPB_DEBUG=1 $cfg timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_parity.py tests/test_moments.py tests/test_collisions.py tests/test_thermal_bc.py -m gpu -q -x -k "not 3" > gpurun_out/r2_call27_pytest.log 2>&1
  echo "$cfg: $(tail -1 gpurun_out/r2_call27_pytest.log | cut -c1-200)"
done
run() {  # name, env...
  name=$1; shift
  env EPB_DEBUG=1 "$@" timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-parity-check --no-e2e-full \
    > gpurun_out/r2_call27_$name.json 2> gpurun_out/r2_call27_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_call27_$name.json").read().strip().splitlines()[-1])
    print("$name", d["ms_per_step"], d["value"], "e2e/value", d["e2e"]["value"]/d["value"], "kernel", d["roofline"]["kernel_ms"], d["roofline"]["frac"], "mixed", d["mixed_state"]["ms_per_step"], d["mixed_state"]["kernel_ms"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2_call27_$name.err").read()[-1500:])
PY
}
run default EPB_X=1
run prefetch EPB_SLOTS_PREFETCH=1
run cty4 EPB_SLOTS_CTY=4
run default_again EPB_X=1
