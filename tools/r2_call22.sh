#!/bin/bash
# round 2, GPU call 22: where the e2e loop's extra milliseconds go (dump on / off / synchronous), HC in the slot kernels, default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_hc_push.py tests/test_zz_reference_binary_gpu.py -m gpu -q -x > gpurun_out/r2_call22_pytest.log 2>&1; tail -4 gpurun_out/r2_call22_pytest.log | cut -c1-250
for mode in "" "EPB_BENCH_NO_DUMP=1" "EPB_BENCH_SYNC_DUMP=1"; do
  env $mode EPB_BENCH_E2E_BREAKDOWN=1 timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --no-parity-check --no-e2e-full --no-mixed > gpurun_out/r2_call22_bd.json 2> gpurun_out/r2_call22_bd.err
  echo "mode: $mode"; grep "e2e breakdown" gpurun_out/r2_call22_bd.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_call22_bd.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['value']/d['value'])"
done
( time timeout 900 python bench.py > gpurun_out/r2_call22_bench_default.json 2> gpurun_out/r2_call22_bench_default.err ) 2>&1 | grep real
tail -c 3800 gpurun_out/r2_call22_bench_default.json; tail -3 gpurun_out/r2_call22_bench_default.err
