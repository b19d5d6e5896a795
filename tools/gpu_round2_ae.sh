#!/bin/bash
# round-2 GPU call AE: the adopted layout against two more variants (VNDF sampler out of line; skip kernels without the single-step loop)
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for pass in 1 2; do
for lib in libcvr_b200.so libcvr_b200_vndf.so libcvr_b200_fp.so libcvr_b200_vndffp.so; do
  CVR_LIB=$PWD/cudavolumerenderer_b200/$lib timeout 900 python tools/ab_opts.py 1024 32 "" >> gpurun_out/r2ae_ab.log 2>&1
  echo "^^ $lib" >> gpurun_out/r2ae_ab.log
done; done
