#!/bin/bash
# multi-GPU pass: N = number of visible GPUs.  NCCL test, then bench.py under torchrun (weak and strong), then the reference arm.
N=$(python -c "import torch; print(torch.cuda.device_count())")
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/m${N}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/m${N}_pytest.log 2>&1; echo "exit $?" >> gpurun_out/m${N}_pytest.log
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,P2P timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err; echo "exit $?" >> gpurun_out/m${N}_bench.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --no-rows > gpurun_out/m${N}_bench_strong.json 2> gpurun_out/m${N}_bench_strong.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
  bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/m${N}_bench_ref.json 2> gpurun_out/m${N}_bench_ref.err
grep -c "NCCL INFO" gpurun_out/m${N}_bench.err > gpurun_out/m${N}_nccl_lines.txt
echo done
