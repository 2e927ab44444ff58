#!/bin/bash
# ncu capture of the band Cholesky kernel (one launch, --set full with source) at configs[4].
# usage: tools/dev/ncu_band.sh <outdir>
out=${1:-gpurun_out/ncu_band}
mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:band_cholesky -s 2 -c 1 -o $out/band -f \
    python tools/ba_quick.py large 3 > $out/ncu.log 2>&1
ncu -i $out/band.ncu-rep --page raw --csv > $out/band_raw.csv 2>/dev/null
ncu -i $out/band.ncu-rep --page source --csv > $out/band_source.csv 2>/dev/null
ls -la $out
