#!/bin/bash
# round-2 GPU call G: DRAM sectors per random 32-byte gather by load flavour (tools/exp/dram_granule.cu), then the
# whole GPU suite and the bench line on the current build
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
M="dram__sectors_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,l1tex__m_xbar2l1tex_read_sectors.sum,lts__t_sectors_srcunit_ltcfabric.sum,gpu__time_duration.sum"
for g in 0 32; do
  timeout 300 tools/exp/dram_granule 8192 $g > gpurun_out/r2g_granule_plain_$g.log 2>&1; cat gpurun_out/r2g_granule_plain_$g.log
  timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2g_granule_ncu_$g.csv tools/exp/dram_granule 8192 $g > /dev/null 2>&1
done
timeout 300 tools/exp/dram_granule 64 0 > gpurun_out/r2g_granule_plain_l2.log 2>&1; cat gpurun_out/r2g_granule_plain_l2.log
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2g_tests.log; tail -15 gpurun_out/r2g_tests.log | cut -c 1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 1500 gpurun_out/r2g_bench.json; tail -5 gpurun_out/r2g_bench.err
