#!/bin/bash
# round-2 GPU call H: gather ceiling at the C4 footprint + L2-prefetch pipeline (tools/exp/dram_granule.cu),
# A/B of the .L2::64B density load (libcvr_b200_l264.so) on the HBM-resident scenes
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for mib in 8192 34000; do
  timeout 300 tools/exp/dram_granule $mib 0 > gpurun_out/r2h_granule_plain_$mib.log 2>&1; cat gpurun_out/r2h_granule_plain_$mib.log
done
M="dram__sectors_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,l1tex__m_xbar2l1tex_read_sectors.sum,lts__t_sectors_srcunit_ltcfabric.sum,gpu__time_duration.sum,lts__t_requests.sum,lts__t_sectors_lookup_miss.sum"
for f in 14 15 16; do
  timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2h_granule_ncu_$f.csv tools/exp/dram_granule 8192 0 $f > /dev/null 2>&1
done
for lib in libcvr_b200.so libcvr_b200_l264.so libcvr_b200.so libcvr_b200_l264.so; do
  CVR_LIB=$PWD/cudavolumerenderer_b200/$lib CVR_AB_SCENES=hetvol,manix,fbm512,fbm1024,sparse1024 timeout 900 python tools/ab_opts.py 1024 32 "" >> gpurun_out/r2h_ab_l264.log 2>&1
  echo "^^ $lib" >> gpurun_out/r2h_ab_l264.log
done
cat gpurun_out/r2h_ab_l264.log
