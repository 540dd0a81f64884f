#!/bin/bash
# round 2, GPU call 24: CPML on the device vs the oracle (1D/2D/3D, orders, stencils, particles) + the reference's Pukhov deck
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cpml.py tests/test_host_logic.py -m gpu -q > gpurun_out/r2_call24_pytest.log 2>&1; tail -12 gpurun_out/r2_call24_pytest.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_field_solvers.py -m gpu -q -x > gpurun_out/r2_call24_pytest2.log 2>&1; tail -3 gpurun_out/r2_call24_pytest2.log | cut -c1-300
