#!/bin/bash
# round-2 GPU call D: whole GPU suite, fbm sweeps, ncu --set full of hetvol and fbm 1024^3 (C4-like)
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.json
timeout 2400 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log; tail -25 gpurun_out/r2d_tests.log | cut -c 1-300
CVR_AB_SCENES=fbm512,fbm1024,sparse1024 timeout 900 python tools/ab_opts.py 1024 32 "" "warp_slots=96" "track_steps=8" "track_steps=4,track_min_lanes=16" "exit_others=8" "skip=0" "warp_slots=96,track_steps=8" > gpurun_out/r2d_ab_fbm.log 2>&1; cat gpurun_out/r2d_ab_fbm.log
P1="python tools/profile_run.py hetvol 1024 16 2"
$P1 > gpurun_out/r2d_plain1.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_volpt_warp -s 1 -c 1 -f -o gpurun_out/r2d_hetvol $P1 > gpurun_out/r2d_ncu1.log 2>&1
tail -2 gpurun_out/r2d_plain1.log gpurun_out/r2d_ncu1.log
P2="python tools/profile_run.py devfbm:1024 1024 16 2"
$P2 > gpurun_out/r2d_plain2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_volpt_warp -s 1 -c 1 -f -o gpurun_out/r2d_fbm1024 $P2 > gpurun_out/r2d_ncu2.log 2>&1
tail -2 gpurun_out/r2d_plain2.log gpurun_out/r2d_ncu2.log
P3="python tools/profile_run.py hetvol 1024 16 2 rng=philox"
$P3 > gpurun_out/r2d_plain3.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_volpt_warp -s 1 -c 1 -f -o gpurun_out/r2d_hetvol_philox $P3 > gpurun_out/r2d_ncu3.log 2>&1
tail -2 gpurun_out/r2d_plain3.log gpurun_out/r2d_ncu3.log
ls -la gpurun_out/*.ncu-rep
