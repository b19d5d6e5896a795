#!/bin/bash
# usage: tools/ncu_export.sh <name> <cmd...>   (on the GPU box)
# plain run, then ncu --set full of the first k_volpt_warp launch after one warm-up; the report stays in /tmp
# (a .ncu-rep is ~33 MB and gpurun_out/ is capped at 64 MiB), its raw metrics, details and source pages are
# exported as text into gpurun_out/.
name=$1; shift
"$@" > gpurun_out/${name}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_volpt_warp -s 1 -c 1 -f -o /tmp/${name} "$@" > gpurun_out/${name}_ncu.log 2>&1
tail -n 1 gpurun_out/${name}_plain.log
if [ -f /tmp/${name}.ncu-rep ]; then
  ncu -i /tmp/${name}.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${name}.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  ncu -i /tmp/${name}.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
  ls -la gpurun_out/${name}_*
fi
