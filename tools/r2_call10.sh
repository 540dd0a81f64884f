#!/bin/bash
# round 2, GPU call 10 (4 GPUs): four-rank parity (incl. re-cut slabs), C3 with the balancer, C2 and C4 lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -k "four or rebalance" > gpurun_out/r2_call10_pytest_multi.log 2>&1
tail -12 gpurun_out/r2_call10_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 4 --workload c3 --cells 2048 --steps 8 --warmup 3 > gpurun_out/r2_call10_c3_4gpu.json 2> gpurun_out/r2_call10_c3_4gpu.err
tail -c 2600 gpurun_out/r2_call10_c3_4gpu.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_call10_c3_4gpu.err | tail -5
timeout 900 $TR --master-port 29522 bench.py --gpus 4 --steps 8 --warmup 3 --no-parity-check > gpurun_out/r2_call10_c2_4gpu.json 2> gpurun_out/r2_call10_c2_4gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_call10_c2_4gpu.json').read().strip().splitlines()[-1]); print('c2 4gpu:', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel_ms'])"
timeout 900 $TR --master-port 29523 bench.py --gpus 4 --workload c4 --steps 6 --warmup 3 > gpurun_out/r2_call10_c4_4gpu.json 2> gpurun_out/r2_call10_c4_4gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_call10_c4_4gpu.json').read().strip().splitlines()[-1]); print('c4 4gpu:', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d.get('parity_check'))"
