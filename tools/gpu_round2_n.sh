#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cloud.py -q -k "voxel or fusion" > gpurun_out/n_pytest.log 2>&1; echo "exit $?" >> gpurun_out/n_pytest.log
bash tools/k4_sweep.sh > gpurun_out/n_k4_sweep.log 2>&1
