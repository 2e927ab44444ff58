#!/bin/bash
# usage: ab2.sh "ENV1=.. ENV2=.." "ENV..." ...   one bench line (matching only) per environment setting
mkdir -p gpurun_out
i=0
for envs in "$@"; do
i=$((i+1))
env $envs timeout 300 python bench.py --steps 5 --warmup 3 --no-ba --cpu-pairs 0 > gpurun_out/ab_$i.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/ab_$i.log") if x.startswith("{")]
if not l: print("$envs", open("gpurun_out/ab_$i.log").read()[-1500:])
else:
    d=json.loads(l[-1])
    print("[$envs]", round(d["ms_per_step"],2), "%.3e"%d["value"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"], d["clocks"]["reasons"], {k:round(x,2) for k,x in d["kernels_ms_per_step"].items()}, "e2e", round(d["e2e"]["ms_per_step"],2), d["config"]["matches_per_step"])
PY
done
