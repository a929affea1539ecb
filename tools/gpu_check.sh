#!/bin/bash
# One gpurun call: new-kernel parity first, then the encoder A/B, then the whole GPU suite.  Logs under gpurun_out/.
mkdir -p gpurun_out
T=${1:-r02b}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${T}_smi.log 2>&1
timeout 600 python -m pytest tests/test_gpu_ops_parity.py -q -m gpu -k "attention or linear" > gpurun_out/${T}_ops.log 2>&1
echo "ops rc=$?" >> gpurun_out/${T}_ops.log
tail -15 gpurun_out/${T}_ops.log
timeout 600 python tools/encoder_ab.py 60 2 > gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/${T}_full.log 2>&1
echo "full rc=$?" >> gpurun_out/${T}_full.log
tail -15 gpurun_out/${T}_full.log
