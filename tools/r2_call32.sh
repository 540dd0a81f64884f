#!/bin/bash
# round 2, GPU call 32 (2 GPUs): rebalance with the re-agreed exchange capacities; C3 A/B: exchange with / without host round trips
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -k "rebalance" > gpurun_out/r2_call32_pytest_multi.log 2>&1
tail -3 gpurun_out/r2_call32_pytest_multi.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mode in "EPB_X=1" "EPB_EXCHANGE_SYNC=1" "EPB_X=1"; do
  env EPB_DEBUG=1 $mode timeout 900 $TR --master-port 29561 bench.py --gpus 2 --workload c3 --cells 2048 --steps 10 --warmup 3 --no-parity-check > gpurun_out/r2_call32_c3_2gpu.json 2> gpurun_out/r2_call32_c3_2gpu.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_call32_c3_2gpu.json').read().strip().splitlines()[-1]); print('c3 2gpu $mode:', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['c3']['unbalanced_ms_per_step'], d['c3']['balanced_ms_per_step'])"
done
