#!/bin/bash
# 2-GPU box: the sharded encoder (host C + NCCL halo exchange) with the fused long-call path, then bench.py --gpus 2 (reduced sizes)
mkdir -p gpurun_out
T=${1:-r02e}
nvidia-smi -L > gpurun_out/${T}_smi.log 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -m gpu -rxs > gpurun_out/${T}_sharded_test.log 2>&1
echo "sharded test rc=$?" | tee -a gpurun_out/${T}_sharded_test.log
tail -5 gpurun_out/${T}_sharded_test.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 1 --warmup 3 --seconds 120 --sharded-seconds 1800 > gpurun_out/${T}_bench2.log 2> gpurun_out/${T}_bench2.err
echo "bench2 rc=$?"
tail -c 3000 gpurun_out/${T}_bench2.log
tail -5 gpurun_out/${T}_bench2.err
