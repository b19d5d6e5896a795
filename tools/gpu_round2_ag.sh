#!/bin/bash
# round-2 GPU call AG: small-argument sin/cos only (t2), tan only (t1), all three (libcvr_b200.so) against libdevice's functions (notrig)
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for pass in 1 2; do
for lib in libcvr_b200_notrig.so libcvr_b200_t1.so libcvr_b200_t2.so libcvr_b200.so; do
  CVR_LIB=$PWD/cudavolumerenderer_b200/$lib timeout 900 python tools/ab_opts.py 1024 32 "" >> gpurun_out/r2ag_ab.log 2>&1
  echo "^^ $lib" >> gpurun_out/r2ag_ab.log
done; done
python tools/ab_table.py gpurun_out/r2ag_ab.log
