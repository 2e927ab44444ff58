#!/bin/bash
# one matching-only bench line per library build: usage tools/dev/match_ab2.sh <outdir> <lib or ""> ...
out=$1; shift
mkdir -p $out
i=0
for lib in "$@"; do
  i=$((i+1))
  if [ "$lib" = "default" ]; then unset MSFM_B200_LIB; else export MSFM_B200_LIB=$lib; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-ba --cpu-pairs 0 --no-target --no-verify --no-extra > $out/bench_$i.json 2> $out/bench_$i.err
  python - <<PY
import json
l = [x for x in open("$out/bench_$i.json") if x.startswith("{")]
if not l:
    print("$lib", open("$out/bench_$i.err").read()[-800:])
else:
    d = json.loads(l[-1])
    print("$lib", round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["kernels_ms_per_step"].items()}, d["config"].get("matches_this_rank_per_step"))
PY
done
