#!/bin/bash
# round-2 GPU call W2 (8 GPUs): final build (after the code-layout change) -- bench at N = 1, 2, 4, 8, cvr_render --gpus 8
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2w2_bench_n1.json 2> gpurun_out/r2w2_bench_n1.err
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2w2_bench_n$n.json 2> gpurun_out/r2w2_bench_n$n.err
done
timeout 300 cudavolumerenderer_b200/cvr_render synth:manix -k regenerationSK -r 1024 -i 256 --number-of-tiles 10 --interactive 0 --gpus 8 --shard balanced --trials 6 > gpurun_out/r2w2_cli_gpus8.log 2>&1; grep "rendering time\|mean time\|paths per" gpurun_out/r2w2_cli_gpus8.log | tail -8
