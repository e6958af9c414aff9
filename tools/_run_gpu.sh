set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cloud.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_kernels.py > gpurun_out/k4v7_kernels.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_voxel" -c 3 -f -o gpurun_out/k4_v7 python tools/bench_kernels.py --reps 1 > gpurun_out/ncu_k4.log 2>&1
echo done
