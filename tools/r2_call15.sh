#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_call15_pytest.log 2>&1; tail -25 gpurun_out/r2_call15_pytest.log | cut -c1-250
