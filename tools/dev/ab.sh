timeout 600 python -m pytest tests/test_match_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in $VARIANTS; do
MSFM_K1_SINGLE=$v timeout 200 python bench.py --steps 5 --warmup 3 --no-ba --cpu-pairs 0 > gpurun_out/bench_ab$v.log 2>&1
python - <<EOF
import json
l=[x for x in open("gpurun_out/bench_ab$v.log") if x.startswith("{")]
if not l: print(open("gpurun_out/bench_ab$v.log").read()[-2000:])
else:
    d=json.loads(l[-1])
    print("single=$v", round(d["ms_per_step"],2), "%.3e"%d["value"], d["clocks"], {k:round(x,2) for k,x in d["kernels_ms_per_step"].items()}, round(d["e2e"]["ms_per_step"],2), d["config"]["matches_per_step"])
EOF
done
