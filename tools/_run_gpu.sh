set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1
timeout 300 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --frames 4096 > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches_k234.csv python tools/bench_kernels.py --reps 1 > gpurun_out/ncu_launches2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_deproject -s 2 -c 1 -f -o gpurun_out/k1_r1k python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --frames 4096 > gpurun_out/ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_reg|k_voxel|k_transform" -c 12 -f -o gpurun_out/k234_r1 python tools/bench_kernels.py --reps 1 > gpurun_out/ncu_k234.log 2>&1
echo done
