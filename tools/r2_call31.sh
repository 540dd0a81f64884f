#!/bin/bash
# round 2, GPU call 31 (2 GPUs): the exchange without host round trips: two-rank parity, rebalance, bench with breakdown
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -q -k "two or rebalance_two" > gpurun_out/r2_call31_pytest_multi.log 2>&1
tail -6 gpurun_out/r2_call31_pytest_multi.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
EPB_BENCH_E2E_BREAKDOWN=1 timeout 900 $TR --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 --no-mixed > gpurun_out/r2_call31_c2_2gpu.json 2> gpurun_out/r2_call31_c2_2gpu.err
grep "e2e breakdown" gpurun_out/r2_call31_c2_2gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_call31_c2_2gpu.json').read().strip().splitlines()[-1]); print('c2 2gpu:', d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['value']/d['value'], d['roofline']['kernel_ms'], d.get('parity_check'))"
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_call31_c2_2gpu.err | tail -3
timeout 900 $TR --master-port 29552 bench.py --gpus 2 --workload c3 --cells 2048 --steps 8 --warmup 3 > gpurun_out/r2_call31_c3_2gpu.json 2> gpurun_out/r2_call31_c3_2gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_call31_c3_2gpu.json').read().strip().splitlines()[-1]); print('c3 2gpu:', d['ms_per_step'], d['value'], d['c3']['unbalanced_ms_per_step'], d['c3']['balanced_ms_per_step'], d.get('parity_check'))"
