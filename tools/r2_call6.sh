#!/bin/bash
# round 2, GPU call 6: row-block arena layout -- parity subset + bench (compare with call 5 on the same code minus the layout)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_parity.py tests/test_moments.py tests/test_host_replay.py tests/test_zz_reference_binary_gpu.py -m gpu -x -q > gpurun_out/r2_call6_pytest.log 2>&1
tail -6 gpurun_out/r2_call6_pytest.log
for cfg in "0" "1" "0"; do
  EPB_LOAD_MIXED=$cfg timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-parity-check \
    > gpurun_out/r2_call6_bench_mix$cfg.json 2> gpurun_out/r2_call6_bench_mix$cfg.err
  echo "mixed=$cfg"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_call6_bench_mix$cfg.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2_call6_bench_mix$cfg.err").read()[-1500:])
PY
done
EPB_PUSH_VARIANT=3 timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-parity-check > gpurun_out/r2_call6_bench_v3.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r2_call6_bench_v3.json').read().strip().splitlines()[-1]); print('variant3 (round-1 kernel) on this box:', d['ms_per_step'], d['roofline']['kernel_ms'])"
