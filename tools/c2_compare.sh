#!/bin/bash
# full-size C2 (4096^2 x 64 ppc) comparison of the two 2D push kernels, fresh (EPOCH loader: exactly
# ppc per cell) and mixed (Poisson counts) initial states
mkdir -p gpurun_out
for mixed in 0 1; do
  for cfg in "0 8" "3 3"; do
    set -- $cfg
    EPB_DEBUG=1 EPB_LOAD_MIXED=$mixed EPB_PUSH_VARIANT=$1 timeout 900 python bench.py --steps 12 --warmup 3 --sort-interval $2 --no-cpu-baseline 2>gpurun_out/c2cmp.err | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mixed=$mixed variant $1 sort $2: push_ms %.3f step_ms %.3f value %.4e e2e %.4e frac %.3f'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac']))" | tee -a gpurun_out/c2_compare.log
    tail -1 gpurun_out/c2cmp.err
  done
done
