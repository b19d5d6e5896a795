#!/bin/bash
# round-2 GPU call C5 (8 GPUs): strong scaling of C5 (sparse 2048^3, 4096^2 x 16 spp, 8 x 8 tiles) on the final build, N = 1 and 8
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 400 python tools/bench_multi.py c5 > gpurun_out/r2_c5_n1.jsonl 2> gpurun_out/r2_c5_n1.err; cat gpurun_out/r2_c5_n1.jsonl | cut -c 1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/bench_multi.py c5 > gpurun_out/r2_c5_n8.jsonl 2> gpurun_out/r2_c5_n8.err; cat gpurun_out/r2_c5_n8.jsonl | cut -c 1-300; tail -2 gpurun_out/r2_c5_n8.err | cut -c 1-200
