#!/bin/bash
# round-2 GPU call E: failing tests again, ncu exports (hetvol xorwow / philox, fbm 1024^3, manix tiles)
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "near_tie or names or display" > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2e_tests.log; tail -60 gpurun_out/r2e_tests.log | cut -c 1-400
tools/ncu_export.sh r2e_hetvol python tools/profile_run.py hetvol 1024 16 2
tools/ncu_export.sh r2e_fbm1024 python tools/profile_run.py devfbm:1024 1024 16 2
tools/ncu_export.sh r2e_hetvol_philox python tools/profile_run.py hetvol 1024 16 2 rng=philox
tools/ncu_export.sh r2e_manix python tools/profile_run.py manix 1024 16 2
du -sh gpurun_out
