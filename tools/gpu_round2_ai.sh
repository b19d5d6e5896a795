#!/bin/bash
# round-2 GPU call AI: pair-loop counters derived from the generator's Weyl counter and the loop's votes (no per-lane counter in the loop)
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r2ai_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2ai_tests.log; tail -6 gpurun_out/r2ai_tests.log | cut -c 1-300
for pass in 1 2; do
for lib in libcvr_b200_prev.so libcvr_b200.so; do
  CVR_LIB=$PWD/cudavolumerenderer_b200/$lib timeout 900 python tools/ab_opts.py 1024 32 "" >> gpurun_out/r2ai_ab.log 2>&1
  echo "^^ $lib" >> gpurun_out/r2ai_ab.log
done; done
python tools/ab_table.py gpurun_out/r2ai_ab.log
grep "fbm1024\|hetvol" gpurun_out/r2ai_ab.log | head -4
