set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deproject.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/pre1_c2048.log 2>&1
timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 --chunk 1024 > gpurun_out/pre1_c1024.log 2>&1
timeout 300 python tools/bench_kernels.py > gpurun_out/pre1_kernels.log 2>&1
RV_NVCC_EXTRA="-DRV_K1_PRELOAD=0" python -m repas_vision_b200._build --force > gpurun_out/build_tmp.log 2>&1 || echo BUILD FAILED
timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/pre0_c2048.log 2>&1
timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 --chunk 1024 > gpurun_out/pre0_c1024.log 2>&1
echo done
