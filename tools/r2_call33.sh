#!/bin/bash
# round 2, GPU call 33 (8 GPUs), final library: eight-rank 3D parity over the host-round-trip-free exchange, C4 (768^3) and C2 lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -k "eight" > gpurun_out/r2_call33_pytest_multi.log 2>&1
tail -3 gpurun_out/r2_call33_pytest_multi.log | cut -c1-200
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29571 bench.py --gpus 8 --workload c4 --steps 6 --warmup 3 > gpurun_out/r2_call33_c4_8gpu.json 2> gpurun_out/r2_call33_c4_8gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_call33_c4_8gpu.json').read().strip().splitlines()[-1]); print('c4 8gpu:', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['e2e']['value']/d['value'], 'kernel', d['roofline']['kernel_ms'], d.get('parity_check'), 'mixed', d.get('mixed_state',{}).get('ms_per_step'))"
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_call33_c4_8gpu.err | tail -3
timeout 900 $TR --master-port 29572 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_call33_c2_8gpu.json 2> gpurun_out/r2_call33_c2_8gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_call33_c2_8gpu.json').read().strip().splitlines()[-1]); print('c2 8gpu:', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['e2e']['value']/d['value'], 'kernel', d['roofline']['kernel_ms'], d.get('parity_check'), 'mixed', d.get('mixed_state',{}).get('ms_per_step'))"
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_call33_c2_8gpu.err | tail -3
