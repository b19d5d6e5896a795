#!/bin/bash
# round-2 GPU call I: two-sided fetch-skip table (majorant + minorant byte per brick) -- A/B on the usual scenes,
# the whole GPU suite, the bench line
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
CVR_AB_SCENES=hetvol,manix,fbm512,fbm1024,sparse1024 timeout 900 python tools/ab_opts.py 1024 32 "" "rng=philox" > gpurun_out/r2i_ab.log 2>&1; cat gpurun_out/r2i_ab.log
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2i_tests.log; tail -15 gpurun_out/r2i_tests.log | cut -c 1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -c 1500 gpurun_out/r2i_bench.json; tail -5 gpurun_out/r2i_bench.err
