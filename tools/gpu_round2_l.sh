#!/bin/bash
# round-2 GPU call L: 32 warps x 64 registers (CVR_WSKIP_BLOCK=1024) against 28 x 72 on the HBM-resident scenes
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
export CVR_AB_SCENES=manix,fbm512,fbm1024,sparse1024
for lib in libcvr_b200.so libcvr_b200_w1024.so; do
  CVR_LIB=$PWD/cudavolumerenderer_b200/$lib timeout 900 python tools/ab_opts.py 1024 32 "" "rng=philox" >> gpurun_out/r2l_ab.log 2>&1
  echo "^^ $lib" >> gpurun_out/r2l_ab.log
done
cat gpurun_out/r2l_ab.log
