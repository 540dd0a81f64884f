#!/bin/bash
# round 2, GPU call 16 (1 GPU): thermal-wall tests after the test resize
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_thermal_bc.py tests/test_host_logic.py -q > gpurun_out/r2_call16_pytest.log 2>&1; tail -15 gpurun_out/r2_call16_pytest.log | cut -c1-250
