#!/bin/bash
# round-2 GPU call C: whole GPU suite, philox 10 vs 7 rounds A/B
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.json
timeout 2400 python -m pytest tests -q -m gpu -x --durations=8 > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log; tail -25 gpurun_out/r2c_tests.log | cut -c 1-400
CVR_AB_SCENES=bucky,hetvol,manix,fbm512 timeout 600 python tools/ab_opts.py 1024 32 "" "rng=philox" > gpurun_out/r2c_ab_p10.log 2>&1; cat gpurun_out/r2c_ab_p10.log
CVR_LIB=$PWD/cudavolumerenderer_b200/libcvr_b200_p7.so CVR_AB_SCENES=bucky,hetvol,manix,fbm512 timeout 600 python tools/ab_opts.py 1024 32 "" "rng=philox" > gpurun_out/r2c_ab_p7.log 2>&1; cat gpurun_out/r2c_ab_p7.log
