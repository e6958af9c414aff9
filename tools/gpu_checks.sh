#!/bin/bash
# Everything measured on a B200 box for profiles/: parity tests, smoke, bench lines, ncu launch lists and --set full
# captures.  Run from the repo root (e.g. `gpurun --timeout 2400 -- 'bash tools/gpu_checks.sh'`); writes gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
(while true; do nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv,noheader; sleep 0.5; done) > gpurun_out/clocks.csv &
CLK=$!
timeout 600 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
kill $CLK
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1
timeout 300 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1
# ncu passes (never a bench value): launch lists, then one full capture per kernel family
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --frames 4096 > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches_kernels.csv \
  python tools/bench_kernels.py --reps 1 > gpurun_out/ncu_launches2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_deproject -s 2 -c 1 -f -o gpurun_out/k1 \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --frames 4096 > gpurun_out/ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reg -c 3 -f -o gpurun_out/k2 \
  python tools/bench_kernels.py --reps 1 > gpurun_out/ncu_k2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_voxel|k_transform" -c 8 -f -o gpurun_out/k34 \
  python tools/bench_kernels.py --reps 1 > gpurun_out/ncu_k34.log 2>&1
echo done
