#!/bin/bash
# round 2, GPU call 35: CPML tests after the fng fix (device laser plane for the Lehe solvers)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cpml.py tests/test_field_solvers.py -m gpu -q > gpurun_out/r2_call35_pytest.log 2>&1; tail -4 gpurun_out/r2_call35_pytest.log | cut -c1-250
