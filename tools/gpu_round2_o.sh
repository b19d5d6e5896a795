#!/bin/bash
# round-2 GPU call O (8 GPUs): device-group fixes -- alpha after the cross-device sum, communicators cached across groups
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "group or shard or gpus or sharding" > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2o_tests.log; tail -5 gpurun_out/r2o_tests.log | cut -c 1-300
for g in 8 4 2; do
timeout 300 cudavolumerenderer_b200/cvr_render synth:manix -k regenerationSK -r 1024 -i 256 --number-of-tiles 10 --interactive 0 --gpus $g --shard balanced --trials 6 > gpurun_out/r2o_cli_gpus$g.log 2>&1; tail -3 gpurun_out/r2o_cli_gpus$g.log
done
timeout 300 cudavolumerenderer_b200/cvr_render synth:manix -k regenerationSK -r 1024 -i 256 --number-of-tiles 10 --interactive 0 --trials 6 > gpurun_out/r2o_cli_gpus1.log 2>&1; tail -3 gpurun_out/r2o_cli_gpus1.log
