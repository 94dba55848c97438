#!/bin/bash
# K1 (work-stealing version): launch list of the default bench command + one --set full capture of exact_scan_kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --queries-per-step 8 --no-cpu-baseline > gpurun_out/e12_plain.json 2>/dev/null; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/e12_k1_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/e12_k1_launches.log 2>&1; echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:exact_scan_kernel -s 4 -c 1 -o gpurun_out/prof_k1_v2 \
    python bench.py --steps 3 --warmup 3 --queries-per-step 8 --no-cpu-baseline --no-e2e > gpurun_out/e12_k1_full.log 2>&1; echo "full rc=$?"
ncu --set full --clock-control none -k regex:exact_scan_kernel -s 4 -c 1 -o gpurun_out/prof_k1_v2_q1 \
    python bench.py --steps 3 --warmup 3 --queries-per-step 1 --no-cpu-baseline --no-e2e > gpurun_out/e12_k1_full_q1.log 2>&1; echo "full q1 rc=$?"
