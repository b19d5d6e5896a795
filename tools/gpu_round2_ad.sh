#!/bin/bash
# round-2 GPU call AD: which cold code to move out of line (kernel 54 KB against a 32 KB L1.5 instruction cache)
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for pass in 1 2; do
for lib in libcvr_b200.so libcvr_b200_cold.so libcvr_b200_b.so libcvr_b200_c.so libcvr_b200_d.so libcvr_b200_f.so libcvr_b200_g.so; do
  CVR_LIB=$PWD/cudavolumerenderer_b200/$lib timeout 900 python tools/ab_opts.py 1024 32 "" >> gpurun_out/r2ad_ab.log 2>&1
  echo "^^ $lib" >> gpurun_out/r2ad_ab.log
done; done
cat gpurun_out/r2ad_ab.log
