#!/bin/bash
# round 2, GPU call 42: moment tests on extents that are no multiple of the tile (padding columns of k_moment2_slots)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_moments.py -m gpu -q > gpurun_out/r2_call42_pytest.log 2>&1; tail -3 gpurun_out/r2_call42_pytest.log | cut -c1-300
