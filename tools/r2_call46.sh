#!/bin/bash
# round 2, GPU call 46: cost of one window shift next to one step, 1024^2 x 64 ppc on one GPU
mkdir -p gpurun_out
timeout 40 python tools/window_shift_time.py 1024 > gpurun_out/r2_call46_window_shift_time.json 2> gpurun_out/r2_call46.err; cat gpurun_out/r2_call46_window_shift_time.json; tail -3 gpurun_out/r2_call46.err
