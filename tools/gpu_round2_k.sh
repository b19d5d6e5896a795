#!/bin/bash
# round-2 GPU call K: one- vs two-sided skip table, table level vs L1 size, against the previous build
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
export CVR_AB_SCENES=manix,fbm512,fbm1024,sparse1024
CVR_LIB=$PWD/cudavolumerenderer_b200/libcvr_b200_head.so timeout 900 python tools/ab_opts.py 1024 32 "" > gpurun_out/r2k_ab.log 2>&1
echo "^^ previous build" >> gpurun_out/r2k_ab.log
timeout 1800 python tools/ab_opts.py 1024 32 "" "skip_sides=2" "skip=16" "skip=16,skip_sides=2" "skip=32" "skip=32,skip_sides=2" "skip=64" "skip=64,skip_sides=2" "rng=philox" "rng=philox,skip_sides=2" >> gpurun_out/r2k_ab.log 2>&1
cat gpurun_out/r2k_ab.log
