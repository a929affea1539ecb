#!/bin/bash
# The gpurun recipes of this round in one place (each stage is what one call ran; logs under gpurun_out/):
#   tests      whole GPU suite, then the encoder A/B on the 60 s and 10-minute clips
#   profile    launch list of the 10-minute encoder pass + ncu --set full of the encoder kernels on the 60 s clip
#   sharded    (2 GPUs) tests/test_gpu_sharded.py + bench.py --gpus 2 with reduced sizes
#   final      tools/gpu_final.sh (smoke, reference arm, bench)
mkdir -p gpurun_out
S=${1:-tests}; T=${2:-r02}
case $S in
tests)
  timeout 1500 python -m pytest tests -q -m gpu -rxs > gpurun_out/${T}_full.log 2>&1; echo "full rc=$?" >> gpurun_out/${T}_full.log; tail -8 gpurun_out/${T}_full.log
  timeout 600 python tools/encoder_ab.py 60 2 > gpurun_out/${T}_ab60.log 2>&1; cat gpurun_out/${T}_ab60.log
  timeout 600 python tools/encoder_ab.py 600 2 0,4 > gpurun_out/${T}_ab600.log 2>&1; cat gpurun_out/${T}_ab600.log ;;
profile)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/${T}_launches600.csv python tools/profile_run.py 600 1 > gpurun_out/${T}_launches600.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_attn_tc|k_gemm_tc2|k_vt_planes|k_rmsnorm_rows_planes|k_rope_table|k_split_planes" -s 40 -c 12 -o gpurun_out/${T}_encoder -f python tools/profile_run.py 60 1 > gpurun_out/${T}_ncu.log 2>&1
  ls -la gpurun_out | tail -4 ;;
sharded)
  timeout 900 python -m pytest tests/test_gpu_sharded.py -q -m gpu -rxs > gpurun_out/${T}_sharded_test.log 2>&1; tail -4 gpurun_out/${T}_sharded_test.log
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 1 --warmup 3 --seconds 120 --sharded-seconds 1800 > gpurun_out/${T}_bench2.log 2> gpurun_out/${T}_bench2.err
  tail -c 2500 gpurun_out/${T}_bench2.log ;;
final) bash tools/gpu_final.sh $T ;;
esac
