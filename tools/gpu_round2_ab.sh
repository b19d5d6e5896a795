#!/bin/bash
# round-2 GPU call AB: regeneration order (rows vs 8x4 pixel blocks) -- the 2-D analogue of the reference's Morton-ordered regeneration
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python tools/ab_opts.py 1024 32 "" "regen_order=block" "" "regen_order=block" > gpurun_out/r2ab_regen_order.log 2>&1
cat gpurun_out/r2ab_regen_order.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "regen_order" > gpurun_out/r2ab_tests.log 2>&1; tail -5 gpurun_out/r2ab_tests.log | cut -c 1-200
