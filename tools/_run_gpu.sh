set -x
mkdir -p gpurun_out
for cw in 16; do
  RV_NVCC_EXTRA="-DRV_K1_CW=$cw" python -m repas_vision_b200._build --force > gpurun_out/build_tmp.log 2>&1 || echo BUILD FAILED $cw
  timeout 600 python -m pytest tests/test_gpu_deproject.py -m gpu -x -q > gpurun_out/pytest_cw$cw.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_cw$cw.log
  timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/cw${cw}_c2048.log 2>&1
  timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 --chunk 1024 > gpurun_out/cw${cw}_c1024.log 2>&1
  timeout 300 python tools/bench_kernels.py > gpurun_out/cw${cw}_kernels.log 2>&1
done
echo done
