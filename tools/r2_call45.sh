#!/bin/bash
# round 2, GPU call 45: the balancer's redistribution after the refactor that the window shares (single rank), window tests again
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_multi.py tests/test_window.py -m gpu -k "rebalance_single or window_matches" -q > gpurun_out/r2_call45_pytest.log 2>&1; tail -5 gpurun_out/r2_call45_pytest.log | cut -c1-300
