#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02g}
timeout 900 python -m pytest tests -q -m gpu -rxs > gpurun_out/${T}_full.log 2>&1
echo "full rc=$?" >> gpurun_out/${T}_full.log
tail -8 gpurun_out/${T}_full.log
timeout 600 python tools/encoder_ab.py 60 2 0,4 > gpurun_out/${T}_ab60.log 2>&1
cat gpurun_out/${T}_ab60.log
timeout 600 python tools/encoder_ab.py 600 2 0 > gpurun_out/${T}_ab600.log 2>&1
cat gpurun_out/${T}_ab600.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_attn_tc|k_gemm_tc2" -s 12 -c 5 -o gpurun_out/${T}_encoder -f python tools/profile_run.py 60 1 > gpurun_out/${T}_ncu.log 2>&1
tail -1 gpurun_out/${T}_ncu.log
