set -x
mkdir -p gpurun_out
for cfg in "0 4" "0 16" "1 4" "0 2"; do
  set -- $cfg
  RV_NVCC_EXTRA="-DRV_K1_TWO_PHASE=$1 -DRV_K1_TICKET_BATCH=$2" python -m repas_vision_b200._build --force > gpurun_out/build_tmp.log 2>&1 || echo BUILD FAILED $cfg
  timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/tb_p$1_b$2_c2048.log 2>&1
  timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 5 --chunk 1024 > gpurun_out/tb_p$1_b$2_c1024.log 2>&1
done
echo done
