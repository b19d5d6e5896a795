#!/bin/bash
# round-2 GPU call F: L2 fetch granularity A/B (gather microbenchmark + HBM-resident scenes), parity tests, new bench line
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python tools/gather_sweep.py > gpurun_out/r2f_gather.log 2>&1; cat gpurun_out/r2f_gather.log
CVR_AB_SCENES=hetvol,manix,fbm512,fbm1024,sparse1024 timeout 900 python tools/ab_opts.py 1024 32 "" "l2_fetch=32" "l2_fetch=64" "l2_fetch=128" "l2_fetch=default" > gpurun_out/r2f_ab.log 2>&1; cat gpurun_out/r2f_ab.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "near_tie or names or display" > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log; tail -30 gpurun_out/r2f_tests.log | cut -c 1-300
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 6000 gpurun_out/r2f_bench.json; tail -5 gpurun_out/r2f_bench.err
