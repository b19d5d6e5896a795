#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python tools/diag_divergence.py > gpurun_out/diag.log 2>&1; tail -80 gpurun_out/diag.log
