#!/bin/bash
# round-2 GPU call N (8 GPUs): bench at N = 1, 2, 4, 8 (weak scaling of the hetvol step + strong scaling of C3 with the balanced
# plan and one reduce to rank 0), the device-group / shard-plan / cvr_render --gpus tests with 8 devices visible
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests -q -m gpu -x -k "group or shard or gpus or sharding" > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2n_tests.log; tail -5 gpurun_out/r2n_tests.log | cut -c 1-300
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2n_bench_n$n.json 2> gpurun_out/r2n_bench_n$n.err
  tail -c 600 gpurun_out/r2n_bench_n$n.json; tail -3 gpurun_out/r2n_bench_n$n.err
done
timeout 300 cudavolumerenderer_b200/cvr_render synth:manix -k regenerationSK -r 1024 -i 256 --number-of-tiles 10 --interactive 0 --gpus 8 --shard balanced --trials 5 > gpurun_out/r2n_cli_gpus8.log 2>&1; tail -8 gpurun_out/r2n_cli_gpus8.log
timeout 300 cudavolumerenderer_b200/cvr_render synth:manix -k regenerationSK -r 1024 -i 256 --number-of-tiles 10 --interactive 0 --trials 5 > gpurun_out/r2n_cli_gpus1.log 2>&1; tail -4 gpurun_out/r2n_cli_gpus1.log
