#!/bin/bash
# round-2 GPU call A: the whole GPU test suite (parity stats land in gpurun_out/parity_stats.json), the bench line, C1-C5 table
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.json
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt; nproc >> gpurun_out/a_gpu.txt; free -g >> gpurun_out/a_gpu.txt
timeout 2400 python -m pytest tests -q -m gpu --durations=15 > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/a_tests.log
tail -40 gpurun_out/a_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; tail -c 1500 gpurun_out/a_bench.json
timeout 900 python tools/bench_configs.py > gpurun_out/a_configs.jsonl 2> gpurun_out/a_configs.err; cut -c 1-400 gpurun_out/a_configs.jsonl
