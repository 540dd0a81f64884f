#!/bin/bash
# round 2, GPU call 43: moving window on the device (epb_shift_window) against the oracle, 1D / 2D / 3D
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_window.py -m gpu -q > gpurun_out/r2_call43_pytest.log 2>&1; tail -25 gpurun_out/r2_call43_pytest.log | cut -c1-400
