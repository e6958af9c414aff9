set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/v6_c2048.log 2>gpurun_out/v6_c2048.err
timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 --chunk 1024 > gpurun_out/v6_c1024.log 2>&1
timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 --mode dense_zero --chunk 256 > gpurun_out/v6_dense.log 2>&1
timeout 300 python tools/bench_kernels.py > gpurun_out/v6_kernels.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_deproject -s 2 -c 1 -f -o gpurun_out/k1_v6 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --frames 2048 > gpurun_out/ncu_v6.log 2>&1
echo done
