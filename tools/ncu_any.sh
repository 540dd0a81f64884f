#!/bin/bash
# usage: KREGEX=... tools/ncu_any.sh  (env EPB_* pass through)
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-1} \
  -f -o gpurun_out/prof_${KNAME:-any} python bench.py ${BENCH_ARGS:---cells 2048} --steps ${STEPS:-2} --warmup 3 --no-cpu-baseline > gpurun_out/ncu_any.log 2>&1
echo "ncu rc=$?"
