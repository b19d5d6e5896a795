#!/bin/bash
# one GPU call for a kernel experiment: quick GPU tests, A/B table over the usual scenes (tools/ab_skip.py),
# one ncu --set full capture (hetvol 512^2 x 16) and a short bench run.  usage: gpurun -- bash tools/gpu_ab_round.sh
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q -m gpu > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ab_tests.log
tail -5 gpurun_out/ab_tests.log
timeout 600 python tools/ab_skip.py 1024 16 > gpurun_out/ab_table.log 2>&1; echo "ab rc=$?" >> gpurun_out/ab_table.log
cat gpurun_out/ab_table.log
for s in 0; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_volpt_warp -s 1 -c 1 -f -o gpurun_out/ab_hetvol_skip$s \
    python tools/profile_run.py hetvol 512 16 2 skip=$s > gpurun_out/ab_ncu_skip$s.log 2>&1
  tail -2 gpurun_out/ab_ncu_skip$s.log
done
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err; tail -c 3000 gpurun_out/ab_bench.json
