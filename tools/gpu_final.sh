#!/bin/bash
# final evidence of the round: smoke(), the two bench arms (reference first, as the driver does), ids of the 10-minute pass for an
# offline comparison with the reference's whole-recording trace
mkdir -p gpurun_out
T=${1:-r02h}
timeout 900 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/${T}_smoke.log
tail -3 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
echo "ref rc=$?"; tail -c 600 gpurun_out/${T}_bench_ref.json
VOX_BENCH_SAVE_IDS=1 timeout 1500 python bench.py > gpurun_out/${T}_bench1.json 2> gpurun_out/${T}_bench1.err
echo "bench rc=$?"; tail -c 6000 gpurun_out/${T}_bench1.json; tail -5 gpurun_out/${T}_bench1.err
