#!/bin/bash
# round-2 GPU call X: the other schedulers on fBm (short segments, constant albedo): is a thread-per-path loop competitive there?
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
CVR_AB_SCENES=fbm512,fbm1024 timeout 1500 python tools/ab_opts.py 1024 16 "" "sched=lane" "sched=lane,loop_threshold=8" "sched=lane,loop_threshold=16" "sched=lane,loop_threshold=24" "sched=queued" "sched=sorted" "skip=0" > gpurun_out/r2x_sched_fbm.log 2>&1
cat gpurun_out/r2x_sched_fbm.log
