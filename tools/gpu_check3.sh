#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02d}
timeout 900 python -m pytest tests/test_gpu_ops_parity.py tests/test_gpu_blocks_parity.py -q -m gpu -rxs -s > gpurun_out/${T}_ops.log 2>&1
echo "ops+blocks rc=$?" | tee -a gpurun_out/${T}_ops.log
grep -i "max abs diff\|passed\|failed\|error" gpurun_out/${T}_ops.log | tail -12
timeout 600 python tools/encoder_ab.py 60 2 > gpurun_out/${T}_ab60.log 2>&1
cat gpurun_out/${T}_ab60.log
timeout 1500 python -m pytest tests -q -m gpu -rxs > gpurun_out/${T}_full.log 2>&1
rc=$?
echo "full rc=$rc" >> gpurun_out/${T}_full.log
tail -12 gpurun_out/${T}_full.log
if [ $rc -ne 0 ] || grep -q xfailed gpurun_out/${T}_full.log; then
  for v in VOX_CUDA_FUSE_QKV=0 VOX_CUDA_FUSE=0 VOX_CUDA_ATTN=simt; do
    env $v timeout 600 python -m pytest tests/test_gpu_stream_parity.py tests/test_gpu_stream_scenarios.py -q -m gpu -rxs > gpurun_out/${T}_stream_$v.log 2>&1
    echo "== $v"; tail -5 gpurun_out/${T}_stream_$v.log
  done
fi
timeout 600 python tools/encoder_ab.py 600 2 0,4 > gpurun_out/${T}_ab600.log 2>&1
cat gpurun_out/${T}_ab600.log
