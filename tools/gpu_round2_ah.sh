#!/bin/bash
# round-2 GPU call AH: the adopted code layout (anisotropic HG out of line; skip-table kernels: counter flush out of line + small-argument
# sin / cos) -- whole GPU suite, the usual scenes, the bench line
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r2ah_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2ah_tests.log; tail -4 gpurun_out/r2ah_tests.log | cut -c 1-300
timeout 900 python tools/ab_opts.py 1024 32 "" "rng=philox" > gpurun_out/r2ah_ab.log 2>&1; cat gpurun_out/r2ah_ab.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2ah_bench.json 2> gpurun_out/r2ah_bench.err; tail -c 200 gpurun_out/r2ah_bench.json; tail -3 gpurun_out/r2ah_bench.err
