#!/bin/bash
# ncu --set full captures with source of the four hot kernels (K2, band Cholesky, K1, resolve), raw + source pages as CSV.
# usage: tools/dev/ncu_all.sh <outdir>
out=${1:-gpurun_out/ncu_all}
mkdir -p $out
tools/dev/ncu_k2.sh $out large > /dev/null 2>&1
tools/dev/ncu_band.sh $out > /dev/null 2>&1
M="python bench.py --steps 1 --warmup 1 --no-ba --cpu-pairs 0 --cpu-data --no-extra --no-target --no-verify"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:match_pair_kernel -s 2 -c 1 -o $out/k1 -f $M > $out/ncu_k1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:resolve_rows -s 2 -c 1 -o $out/resolve -f $M > $out/ncu_resolve.log 2>&1
for k in k1 resolve; do
  ncu -i $out/$k.ncu-rep --page raw --csv > $out/${k}_raw.csv 2>/dev/null
  ncu -i $out/$k.ncu-rep --page source --csv > $out/${k}_source.csv 2>/dev/null
done
rm -f $out/*.ncu-rep
ls -la $out
