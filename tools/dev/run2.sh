#!/bin/bash
mkdir -p gpurun_out
for d in 1 11 12 13; do
echo "=== MSFM_K1_DEBUG=$d"
MSFM_K1_DEBUG_RAW=1 MSFM_K1_DEBUG=$d timeout 300 python tools/prof_run.py match 24 > gpurun_out/dbg_$d.log 2>&1; grep -a "K1" gpurun_out/dbg_$d.log | grep -a -v "raw" | head -2;
done
grep -a "K1 raw" gpurun_out/dbg_1.log | head -16
echo "=== BA big"
timeout 600 python bench.py --steps 2 --warmup 3 --images 24 --cpu-pairs 0 --ba-cams 1329 --ba-pts 542000 --ba-track 9.2 > gpurun_out/bench_ba_big.json 2> gpurun_out/bench_ba_big.err
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_ba_big.json") if x.startswith("{")]
print(json.dumps(json.loads(l[-1])["ba"], indent=1)[:3000] if l else open("gpurun_out/bench_ba_big.err").read()[-2000:])
PY
