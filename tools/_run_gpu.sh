set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_kernels.py > gpurun_out/k2v2_kernels.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reg -c 4 -f -o gpurun_out/k2_v2 python tools/bench_kernels.py --reps 1 > gpurun_out/ncu_k2.log 2>&1
echo done
