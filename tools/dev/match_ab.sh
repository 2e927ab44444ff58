#!/bin/bash
# Matching parity tests + one matching-only bench line with the per-kernel times (development aid).
# usage: tools/dev/match_ab.sh <outdir>
out=${1:-gpurun_out/match_ab}
mkdir -p $out
timeout 400 python -m pytest tests/test_match_gpu.py tests/test_host_cpp.py -q -m gpu -x > $out/pytest.log 2>&1
tail -3 $out/pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-ba --cpu-pairs 0 --no-target --no-verify > $out/bench.json 2> $out/bench.err
python - <<PY
import json
l = [x for x in open("$out/bench.json") if x.startswith("{")]
if not l:
    print(open("$out/bench.err").read()[-1500:])
else:
    d = json.loads(l[-1])
    print(round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["kernels_ms_per_step"].items()}, d["match_stats"],
          {k: (round(v["ms_per_step"], 2), v.get("matches_per_step")) for k, v in d.get("distributions", {}).items()},
          d["config"].get("matches_this_rank_per_step"), "e2e", round(d["e2e"]["ms_per_step"], 2))
PY
