#!/bin/bash
# round-2 GPU call S: where the fixed ~20 ms of a cvr_render trial go (group path on one device, phase timer)
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
CVR_TRACE_SLOW=0.2 timeout 300 cudavolumerenderer_b200/cvr_render synth:manix -k regenerationSK -r 1024 -i 256 --number-of-tiles 10 --interactive 0 --shard balanced --trials 3 > gpurun_out/r2s_cli_group1.log 2>&1; grep -v "^\[config\]" gpurun_out/r2s_cli_group1.log | tail -40
CVR_TRACE_SLOW=0.2 timeout 300 cudavolumerenderer_b200/cvr_render synth:manix -k regenerationSK -r 1024 -i 256 --number-of-tiles 10 --interactive 0 --trials 2 > gpurun_out/r2s_cli_plain1.log 2>&1; grep -v "^\[config\]" gpurun_out/r2s_cli_plain1.log | tail -20
