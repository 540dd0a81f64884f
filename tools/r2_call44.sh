#!/bin/bash
# round 2, GPU call 44 (2 GPUs): moving window across two ranks in x against the multi-rank oracle
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_multi.py -k "two_ranks and window2d" -q -x > gpurun_out/r2_call44_pytest.log 2>&1; tail -30 gpurun_out/r2_call44_pytest.log | cut -c1-600
