#!/bin/bash
# round-2 GPU call B: divergence diagnostic, new parity + philox + group tests, xorwow / philox A/B
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python tools/diag_divergence.py > gpurun_out/r2b_diag.log 2>&1; tail -60 gpurun_out/r2b_diag.log | cut -c 1-600
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "philox or statistical or names or shard or group or fused or sharding" > gpurun_out/r2b_tests.log 2>&1; tail -30 gpurun_out/r2b_tests.log
timeout 900 python tools/ab_opts.py 1024 32 "" "rng=philox" "warp_slots=64" "rng=philox,warp_slots=64" > gpurun_out/r2b_ab.log 2>&1; cat gpurun_out/r2b_ab.log
