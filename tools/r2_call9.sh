#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_collisions.py tests/test_gpu_parity.py tests/test_moments.py -m gpu -q > gpurun_out/r2_call9_pytest.log 2>&1
tail -25 gpurun_out/r2_call9_pytest.log
