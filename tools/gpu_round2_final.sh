#!/bin/bash
# round-2 final GPU call: whole GPU suite, smoke, bench line, ncu captures (hetvol bench launch, manix, fbm 1024^3) and the launch
# list of the bench command -- all on the final build
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log; tail -4 gpurun_out/r2f_tests.log | cut -c 1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; cat gpurun_out/r2f_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 300 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
tools/ncu_export.sh r2f_hetvol python tools/profile_run.py hetvol 1024 64 2
tools/ncu_export.sh r2f_manix python tools/profile_run.py manix 1024 32 2
tools/ncu_export.sh r2f_fbm1024 python tools/profile_run.py devfbm:1024 1024 16 2
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r2f_bench_s2.json 2> gpurun_out/r2f_bench_s2.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2f_ncu_launches.log 2>&1
du -sh gpurun_out
