#!/bin/bash
# evidence for the final encoder kernels: launch list of the 10-minute encoder pass, ncu --set full of the attention / GEMM / small kernels
mkdir -p gpurun_out
T=${1:-r02f}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/${T}_launches600.csv python tools/profile_run.py 600 1 > gpurun_out/${T}_launches600.log 2>&1
tail -1 gpurun_out/${T}_launches600.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_attn_tc|k_gemm_tc2|k_vt_planes|k_rmsnorm_rows_planes|k_rope_table|k_split_planes" -s 40 -c 12 -o gpurun_out/${T}_encoder -f python tools/profile_run.py 60 1 > gpurun_out/${T}_ncu.log 2>&1
tail -1 gpurun_out/${T}_ncu.log
ls -la gpurun_out/ | tail -5
