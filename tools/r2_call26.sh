#!/bin/bash
# round 2, GPU call 26: bisect of the e2e loop's extra 2 ms per step (pure steps / + source planes / + scalars / + dump)
mkdir -p gpurun_out
for mode in "EPB_BENCH_NO_SRC=1 EPB_BENCH_NO_SCAL=1" "EPB_BENCH_NO_SCAL=1" "EPB_BENCH_NO_DUMP=1" "EPB_X=1"; do
  env $mode EPB_BENCH_E2E_BREAKDOWN=1 timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-parity-check --no-e2e-full --no-mixed > gpurun_out/r2_call26_bd.json 2> gpurun_out/r2_call26_bd.err
  echo "mode: $mode"; grep "e2e breakdown" gpurun_out/r2_call26_bd.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_call26_bd.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['value']/d['value'])"
done
