#!/bin/bash
# usage: tools/ncu_capture.sh <name> <kernel-regex> <skip> <cmd...>
# One `ncu --set full` capture of one launch; exports the raw and source pages as CSV into gpurun_out/
# (the .ncu-rep itself is too large to bring back) and removes the report.
name=$1; regex=$2; skip=$3; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o /tmp/$name "$@" > gpurun_out/${name}.ncu.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}.raw.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/${name}.source.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page details --csv > gpurun_out/${name}.details.csv 2>/dev/null
rm -f /tmp/$name.ncu-rep
ls -la gpurun_out/${name}.*
