#!/bin/bash
# Run on the B200 box through gpurun: GPU parity tests, the C2 bench line, the ncu launch
# list of the same bench command (reduced grid) and one full ncu capture of the push kernel.
# Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
STAGES="${1:-tests bench launches ncu}"
for s in $STAGES; do
case $s in
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log ;;
testsv)
  # 2D parity tests under the alternative push kernels (EPB_PUSH_VARIANT is read once per process)
  for v in ${TESTVARIANTS:-3 2 4}; do
    EPB_DEBUG=1 EPB_PUSH_VARIANT=$v timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py -m gpu -q \
      -k "2-n or 2d or sort_interval or open_boundaries or foil or full_size or cuda_path" > gpurun_out/pytest_gpu_v$v.log 2>&1
    echo "variant $v pytest rc=$?" >> gpurun_out/pytest_gpu_v$v.log; tail -4 gpurun_out/pytest_gpu_v$v.log
  done ;;
smoke)
  timeout 600 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log ;;
bench)
  timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err ;;
benchref)
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  cat gpurun_out/bench_ref.json ;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches.csv python bench.py --cells ${CELLS:-2048} --steps 4 --warmup 3 --no-cpu-baseline \
      > gpurun_out/launches_bench.log 2>&1
  echo "launches rc=$?" ;;
ncu)
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:push_${NCU_KERNEL:-tiled} -s ${NCU_SKIP:-3} -c ${NCU_COUNT:-2} \
      -f -o gpurun_out/prof_push python bench.py --cells 2048 --steps 2 --warmup 3 --no-cpu-baseline \
      > gpurun_out/ncu_push.log 2>&1
  echo "ncu rc=$?" ;;
multi)
  timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1
  echo "pytest multi rc=$?" >> gpurun_out/pytest_multi.log; tail -30 gpurun_out/pytest_multi.log ;;
benchN)
  NG=$(nvidia-smi -L | wc -l)
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $NG --cells ${BENCH_N:-2048} --steps 8 --warmup 3 > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err
  echo "benchN rc=$?"; cat gpurun_out/bench_n$NG.json; tail -5 gpurun_out/bench_n$NG.err ;;
variants)
  for v in ${VARIANTS:-0 1 2 3}; do
    for si in ${SORTS:-4}; do
      EPB_DEBUG=1 EPB_PUSH_VARIANT=$v timeout 600 python bench.py --cells 2048 --steps 8 --warmup 4 --sort-interval $si --no-cpu-baseline 2>/dev/null | \
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant $v sort $si: push_ms %.3f step_ms %.3f'%(d['roofline']['kernel_ms'], d['ms_per_step']))" | tee -a gpurun_out/variants.log
    done
  done ;;
experiments)
  for e in ${EXPS:-0 1 2 3 8 11 4 15}; do
    EPB_DEBUG=1 EPB_PUSH_EXPERIMENT=$e timeout 600 python bench.py --cells 2048 --steps 8 --warmup 4 --sort-interval 4 --no-cpu-baseline 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('experiment $e: push_ms %.3f step_ms %.3f'%(d['roofline']['kernel_ms'], d['ms_per_step']))" | tee -a gpurun_out/experiments.log
  done ;;
bench3d)
  for no3d in ${NO3D:-0 1}; do
    EPB_DEBUG=1 EPB_NO_TILED_3D=$no3d timeout 900 python bench.py --workload c4 --cells ${CELLS3D:-192} --steps 4 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench3d.err | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('3d no_tiled=$no3d: push_ms %.3f step_ms %.3f value %.3e frac %.3f'%(d['roofline']['kernel_ms'], d['ms_per_step'], d['value'], d['roofline']['frac']))" | tee -a gpurun_out/bench3d.log
    tail -2 gpurun_out/bench3d.err
  done ;;
micro)
  (cd tools && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu) > gpurun_out/micro_build.log 2>&1
  timeout 300 tools/microbench > gpurun_out/microbench.json 2>&1; cat gpurun_out/microbench.json ;;
esac
done
