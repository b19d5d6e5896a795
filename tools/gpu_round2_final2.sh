#!/bin/bash
# round-2 final GPU call (after the code-layout change): ncu captures (hetvol bench launch, manix, fbm 1024^3), launch list of the bench command
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
tools/ncu_export.sh r2g_hetvol python tools/profile_run.py hetvol 1024 64 2
tools/ncu_export.sh r2g_manix python tools/profile_run.py manix 1024 32 2
tools/ncu_export.sh r2g_fbm1024 python tools/profile_run.py devfbm:1024 1024 16 2
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r2g_bench_s2.json 2> gpurun_out/r2g_bench_s2.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2g_ncu_launches.log 2>&1
du -sh gpurun_out
