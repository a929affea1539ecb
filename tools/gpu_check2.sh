#!/bin/bash
# which fixture shows the near-tie flip, and with which kernel variant; then the launch list and ncu captures of the encoder kernels
mkdir -p gpurun_out
T=${1:-r02c}
for v in default attn_simt gemm_v1; do
  case $v in
    default)   E="" ;;
    attn_simt) E="VOX_CUDA_ATTN=simt" ;;
    gemm_v1)   E="VOX_CUDA_GEMM=v1" ;;
  esac
  env $E timeout 900 python -m pytest tests/test_gpu_stream_parity.py tests/test_gpu_stream_scenarios.py tests/test_gpu_multistream.py tests/test_gpu_verify.py -q -m gpu -rxs > gpurun_out/${T}_stream_$v.log 2>&1
  echo "== $v"; tail -6 gpurun_out/${T}_stream_$v.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches.csv python tools/profile_run.py 60 1 > gpurun_out/${T}_launches.log 2>&1
tail -2 gpurun_out/${T}_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_attn_tc|k_gemm_tc2|k_vt_planes|k_split_planes|k_rope_split|k_rmsnorm_rows" -s 60 -c 14 -o gpurun_out/${T}_encoder -f python tools/profile_run.py 60 1 > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
ls -la gpurun_out/
