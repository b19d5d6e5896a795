#!/bin/bash
# round-2 GPU call AA (2 GPUs): the resolve fused with the cross-device sum over peer memory against the ncclReduce path
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
nvidia-smi topo -m | head -6
timeout 900 python -m pytest tests -q -m gpu -x -k "group or shard or gpus or sharding" > gpurun_out/r2aa_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2aa_tests.log; tail -12 gpurun_out/r2aa_tests.log | cut -c 1-300
for red in peer nccl peer nccl; do
CVR_TRACE_SLOW=0.3 timeout 300 cudavolumerenderer_b200/cvr_render synth:manix -k regenerationSK -r 1024 -i 256 --number-of-tiles 10 --interactive 0 --gpus 2 --shard balanced --trials 6 --option group_reduce=$red > gpurun_out/r2aa_cli_$red.log 2>&1; echo "== $red"; grep "rendering time\|mean time" gpurun_out/r2aa_cli_$red.log | tail -4; grep "cvr_group_render" gpurun_out/r2aa_cli_$red.log | tail -3
done
