#!/bin/bash
# final single-GPU evidence: full GPU test suite, smoke, both bench arms, C1-C5, ncu launch list + full capture
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/f_tests.log
tail -6 gpurun_out/f_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; tail -2 gpurun_out/f_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; cat gpurun_out/f_bench_ref.json | cut -c1-400
timeout 600 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; cat gpurun_out/f_bench.json | cut -c1-300
timeout 900 python tools/bench_configs.py > gpurun_out/f_configs.jsonl 2> gpurun_out/f_configs.err; cut -c1-260 gpurun_out/f_configs.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/f_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_volpt_warp -s 1 -c 1 -f -o gpurun_out/f_bench_warp python bench.py --steps 1 --warmup 1 > gpurun_out/f_ncu_full.log 2>&1; tail -2 gpurun_out/f_ncu_full.log
