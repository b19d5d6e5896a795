#!/bin/bash
# round-2 GPU call V: batch parameters on fBm 1024^3 (the sweeps of round 1 were on fBm 512^3 and hetvol)
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
CVR_AB_SCENES=fbm1024 timeout 1500 python tools/ab_opts.py 1024 32 "" "track_steps=4" "track_steps=8" "track_steps=32" "track_min_lanes=6" "track_min_lanes=8" "track_min_lanes=16" "track_min_lanes=20" "track_min_lanes=24" "exit_others=0" "exit_others=8" "exit_others=24" "exit_others=32" "exit_others=48" "track_steps=8,track_min_lanes=16" "track_steps=8,track_min_lanes=20,exit_others=24" "warp_slots=96" "warp_slots=32" "policy=1" > gpurun_out/r2v_sweep_fbm1024.log 2>&1
cat gpurun_out/r2v_sweep_fbm1024.log
