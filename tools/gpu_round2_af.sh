#!/bin/bash
# round-2 GPU call AF: small-argument trig (libdevice fast paths alone) -- exhaustive equality check, then A/B against libdevice's functions
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "small_angle_trig or reference_kernel_same_seeds or nan_pixels" > gpurun_out/r2af_tests.log 2>&1; tail -5 gpurun_out/r2af_tests.log | cut -c 1-300
for pass in 1 2; do
for lib in libcvr_b200_notrig.so libcvr_b200.so; do
  CVR_LIB=$PWD/cudavolumerenderer_b200/$lib timeout 900 python tools/ab_opts.py 1024 32 "" >> gpurun_out/r2af_ab.log 2>&1
  echo "^^ $lib" >> gpurun_out/r2af_ab.log
done; done
python tools/ab_table.py gpurun_out/r2af_ab.log
