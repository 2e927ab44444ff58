#!/bin/bash
# evidence run of the round's final state: everything lands in gpurun_out/final/
O=gpurun_out/final; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,driver_version --format=csv > $O/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench_final.json 2> $O/bench_final.err; tail -c 600 $O/bench_final.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2>&1; tail -c 400 $O/bench_reference_arm.json; echo
MSFM_K1_SINGLE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-ba --cpu-pairs 0 > $O/bench_single_cta_k1.json 2>&1
MSFM_K1_DEBUG_RAW=1 MSFM_K1_DEBUG=1 timeout 300 python tools/prof_run.py match 24 > $O/k1_phase_timeline.log 2>&1
for d in 11 12 13; do echo "== MSFM_K1_DEBUG=$d" >> $O/k1_phase_timeline.log; MSFM_K1_DEBUG=$d timeout 300 python tools/prof_run.py match 24 2>&1 | grep -a "K1 debug\|K1 timeline" | head -2 >> $O/k1_phase_timeline.log; done
timeout 120 build/microbench_pair > $O/microbench_pair.log 2>&1
timeout 120 build/microbench_alu > $O/microbench_alu.log 2>&1
# ncu: launch list of one bench step, then one full capture of a bench-sized K1 launch and of the BA linearisation kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --no-ba --cpu-pairs 0 --cpu-data > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:match_pair_kernel -s 2 -c 1 -f -o $O/k1_pair python bench.py --steps 1 --warmup 1 --no-ba --cpu-pairs 0 --cpu-data > $O/ncu_k1.log 2>&1
ncu -i $O/k1_pair.ncu-rep --page raw --csv > $O/k1_pair_ncu_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:"point_pass|camera_diag|pair_block" -s 3 -c 3 -f -o $O/k2 python tools/prof_run.py ba > $O/ncu_k2.log 2>&1
ncu -i $O/k2.ncu-rep --page raw --csv > $O/k2_ncu_raw.csv 2>/dev/null
rm -f $O/k2.ncu-rep
ls -la $O
