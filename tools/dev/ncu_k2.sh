#!/bin/bash
# ncu capture of the fused BA linearisation kernel (one launch, --set full with source) + launch list of a short BA run.
# usage: tools/dev/ncu_k2.sh <outdir> [small|large]
out=${1:-gpurun_out/ncu_k2}; which=${2:-small}
mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_linearize -s 3 -c 1 -o $out/k2_$which -f \
    python tools/ba_quick.py $which 3 > $out/ncu_$which.log 2>&1
ncu -i $out/k2_$which.ncu-rep --page raw --csv > $out/k2_${which}_raw.csv 2>/dev/null
ncu -i $out/k2_$which.ncu-rep --page source --csv > $out/k2_${which}_source.csv 2>/dev/null
ls -la $out
