#!/bin/bash
# round 2, GPU call 47: where the time of a window shift goes (EPB_REDIST_TIMING), 1024^2 x 64 ppc
mkdir -p gpurun_out
EPB_DEBUG=1 EPB_REDIST_TIMING=1 timeout 30 python tools/window_shift_time.py 1024 > gpurun_out/r2_call47_window_shift_time.json 2> gpurun_out/r2_call47_phases.txt; cat gpurun_out/r2_call47_window_shift_time.json; grep "epb redistribute" gpurun_out/r2_call47_phases.txt | head -24
