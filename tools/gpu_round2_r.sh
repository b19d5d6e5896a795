#!/bin/bash
# round-2 GPU call R (8 GPUs): cvr_render --gpus N after the fused resolve, with the phase timer; group tests again
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "group or shard or gpus or sharding" > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2r_tests.log; tail -3 gpurun_out/r2r_tests.log | cut -c 1-300
for g in 8 4 2; do
CVR_TRACE_SLOW=0.3 timeout 300 cudavolumerenderer_b200/cvr_render synth:manix -k regenerationSK -r 1024 -i 256 --number-of-tiles 10 --interactive 0 --gpus $g --shard balanced --trials 6 > gpurun_out/r2r_cli_gpus$g.log 2>&1; grep "cvr_group_render\|rendering time\|mean time\|paths per" gpurun_out/r2r_cli_gpus$g.log | tail -14
done
