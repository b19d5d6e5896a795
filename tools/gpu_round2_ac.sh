#!/bin/bash
# round-2 GPU call AC: code-size experiments -- anisotropic HG out of line, CVR_FAST_GGX (VNDF sampler without the acos/atan2/tan round trip)
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for lib in libcvr_b200.so libcvr_b200_cold.so libcvr_b200_ggx.so libcvr_b200_coldggx.so libcvr_b200.so libcvr_b200_cold.so libcvr_b200_ggx.so libcvr_b200_coldggx.so; do
  CVR_LIB=$PWD/cudavolumerenderer_b200/$lib timeout 900 python tools/ab_opts.py 1024 32 "" >> gpurun_out/r2ac_ab.log 2>&1
  echo "^^ $lib" >> gpurun_out/r2ac_ab.log
done
cat gpurun_out/r2ac_ab.log
