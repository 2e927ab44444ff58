#!/bin/bash
# first GPU call of the re-entered session: parity, smoke, A/B of the CTA-pair K1 against the single-CTA K1, full bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
for v in 0 1; do
MSFM_K1_SINGLE=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-ba --cpu-pairs 0 > gpurun_out/bench_ab$v.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_ab$v.log") if x.startswith("{")]
if not l: print(open("gpurun_out/bench_ab$v.log").read()[-2000:])
else:
    d=json.loads(l[-1])
    print("single=$v", round(d["ms_per_step"],2), "%.3e"%d["value"], d["clocks"], {k:round(x,2) for k,x in d["kernels_ms_per_step"].items()}, round(d["e2e"]["ms_per_step"],2), d["config"]["matches_per_step"])
PY
done
MSFM_K1_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-ba --cpu-pairs 0 > gpurun_out/bench_dbg.log 2>&1; grep -a "K1" gpurun_out/bench_dbg.log | tail -6
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json
