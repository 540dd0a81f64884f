#!/bin/bash
# round 2, final evidence on one B200 with the library as committed: full GPU suite, the bench lines, ncu launch list of
# the bench command at its default size, one --set full capture of each hot kernel at the bench size
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2_final_pytest_gpu.txt | cut -c1-200
timeout 900 python bench.py > gpurun_out/r2_final_bench_c2_1gpu.json 2> gpurun_out/r2_final_bench_c2_1gpu.err
timeout 900 python bench.py --workload c4 > gpurun_out/r2_final_bench_c4_1gpu.json 2> gpurun_out/r2_final_bench_c4_1gpu.err
python - <<'PY'
import json
for n in ("c2", "c4"):
    try:
        d = json.loads(open(f"gpurun_out/r2_final_bench_{n}_1gpu.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "kernel", d["roofline"]["kernel_ms"], d["roofline"]["frac"], "mixed", d.get("mixed_state", {}).get("ms_per_step"), d["parity_check"])
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/r2_final_bench_{n}_1gpu.err").read()[-800:])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final_launches_c2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e-full --no-mixed > gpurun_out/r2_final_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:push_slots_2d -s 4 -c 1 -o gpurun_out/r2_final_prof_slots2d -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e-full --no-mixed > gpurun_out/r2_final_prof2d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:push_bag_3d -s 4 -c 1 -o gpurun_out/r2_final_prof_bag3d -f \
  python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check --no-mixed > gpurun_out/r2_final_prof3d.log 2>&1
ls -la gpurun_out/r2_final_*
