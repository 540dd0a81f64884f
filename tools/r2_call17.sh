#!/bin/bash
# round 2, GPU call 17 (8 GPUs): eight-rank parity, C3 with the balancer, C4 (2x2x2 of 384^3) and C2 lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -k "eight" > gpurun_out/r2_call17_pytest_multi.log 2>&1
tail -5 gpurun_out/r2_call17_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29531 bench.py --gpus 8 --workload c3 --cells 2048 --steps 8 --warmup 3 > gpurun_out/r2_call17_c3_8gpu.json 2> gpurun_out/r2_call17_c3_8gpu.err
tail -c 3000 gpurun_out/r2_call17_c3_8gpu.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_call17_c3_8gpu.err | tail -5
timeout 900 $TR --master-port 29533 bench.py --gpus 8 --workload c4 --steps 6 --warmup 3 > gpurun_out/r2_call17_c4_8gpu.json 2> gpurun_out/r2_call17_c4_8gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_call17_c4_8gpu.json').read().strip().splitlines()[-1]); print('c4 8gpu:', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d.get('parity_check'))"
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_call17_c4_8gpu.err | tail -5
timeout 900 $TR --master-port 29532 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r2_call17_c2_8gpu.json 2> gpurun_out/r2_call17_c2_8gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_call17_c2_8gpu.json').read().strip().splitlines()[-1]); print('c2 8gpu:', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d.get('parity_check'))"
timeout 300 python -m pytest tests/test_thermal_bc.py -q 2>&1 | tail -3
