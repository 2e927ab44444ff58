#!/bin/bash
# Evidence run of round 2 on one B200: tests, smoke, both bench arms, ncu launch lists, ncu --set full of K2, sanitizers.
# usage: tools/dev/round2_capture.sh <outdir>     (every step under its own timeout; nothing here is a bench value except bench*.json)
out=${1:-gpurun_out/final_r2}
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $out/gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu --durations=8 > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/steps.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.log 2>&1; echo "smoke rc=$?" >> $out/steps.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?" >> $out/steps.log
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; echo "ref rc=$?" >> $out/steps.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $out/launches_match.csv \
    python bench.py --steps 1 --warmup 1 --no-ba --cpu-pairs 0 --cpu-data --no-extra --no-target --no-verify > $out/ncu_match.log 2>&1; echo "ncu match rc=$?" >> $out/steps.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_ba.csv \
    python tools/ba_quick.py large 3 > $out/ncu_ba.log 2>&1; echo "ncu ba rc=$?" >> $out/steps.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_linearize -s 3 -c 1 -o $out/k2_large -f \
    python tools/ba_quick.py large 3 > $out/ncu_k2.log 2>&1; echo "ncu k2 rc=$?" >> $out/steps.log
ncu -i $out/k2_large.ncu-rep --page raw --csv > $out/k2_large_raw.csv 2>/dev/null
ncu -i $out/k2_large.ncu-rep --page source --csv > $out/k2_large_source.csv 2>/dev/null
timeout 400 compute-sanitizer --tool memcheck python tests/tools/sanitize_run.py > $out/memcheck.log 2>&1; echo "memcheck rc=$?" >> $out/steps.log
timeout 400 compute-sanitizer --tool racecheck python tests/tools/sanitize_run.py > $out/racecheck.log 2>&1; echo "racecheck rc=$?" >> $out/steps.log
cat $out/steps.log
tail -3 $out/pytest_gpu.log
tail -c 400 $out/bench.json
