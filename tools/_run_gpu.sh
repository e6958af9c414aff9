set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1
timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/chk_c2048.log 2>&1
echo done
