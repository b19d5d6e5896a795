#!/bin/bash
# round-2 GPU call J: two-sided fetch-skip table with dummy cells (no verdict live across the load) -- A/B, skip tests
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
CVR_AB_SCENES=hetvol,manix,fbm512,fbm1024,sparse1024 timeout 900 python tools/ab_opts.py 1024 32 "" "rng=philox" > gpurun_out/r2j_ab.log 2>&1; cat gpurun_out/r2j_ab.log
timeout 2400 python -m pytest tests -q -m gpu -x -k "skip or config or c3 or c4 or c5 or statistical" > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2j_tests.log; tail -15 gpurun_out/r2j_tests.log | cut -c 1-300
