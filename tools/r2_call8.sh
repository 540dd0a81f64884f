#!/bin/bash
# round 2, GPU call 8 (2 GPUs): multi-rank parity (incl. the load-balancer remap) and 2-GPU bench lines (C2, C3)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2_call8_pytest_multi.log 2>&1
tail -15 gpurun_out/r2_call8_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_call8_c2_2gpu.json 2> gpurun_out/r2_call8_c2_2gpu.err
tail -c 2500 gpurun_out/r2_call8_c2_2gpu.json; tail -3 gpurun_out/r2_call8_c2_2gpu.err
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --workload c3 --cells 2048 --steps 8 --warmup 3 --no-parity-check > gpurun_out/r2_call8_c3_2gpu.json 2> gpurun_out/r2_call8_c3_2gpu.err
tail -c 2500 gpurun_out/r2_call8_c3_2gpu.json; tail -3 gpurun_out/r2_call8_c3_2gpu.err
