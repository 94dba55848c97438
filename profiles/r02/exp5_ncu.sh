#!/bin/bash
# r02 experiment 5: ncu --set full of the K2 last-segment launch and of the K1 launches (64 queries persistent,
# single query, bf16 rows), raw pages exported as CSV; plus launch lists (gpu__time_duration) of the default bench
cd $GRAFT_REPO_ROOT
NCU=/usr/local/cuda/bin/ncu
B="python bench.py --workload batch_bf16 --no-cpu-baseline --no-e2e --no-parity --steps 1 --warmup 2"
$B > gpurun_out/ncu_k2_plain.json 2> gpurun_out/ncu_k2_plain.err || { echo "plain K2 run failed"; exit 1; }
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:gemm_topk_kernel --launch-skip 17 --launch-count 1 -f -o gpurun_out/k2_r02_last_segment $B > gpurun_out/ncu_k2.log 2>&1
echo "ncu k2 rc=$?"
$NCU -i gpurun_out/k2_r02_last_segment.ncu-rep --page raw --csv > gpurun_out/k2_r02_last_segment_raw.csv 2>/dev/null
python profiles/r02/k1_probe.py > gpurun_out/k1_probe_plain.log 2>&1 || { echo "plain K1 probe failed"; cat gpurun_out/k1_probe_plain.log; exit 1; }
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:exact_scan_kernel --launch-skip 6 --launch-count 3 -f -o gpurun_out/k1_r02_persistent python profiles/r02/k1_probe.py > gpurun_out/ncu_k1.log 2>&1
echo "ncu k1 rc=$?"
$NCU -i gpurun_out/k1_r02_persistent.ncu-rep --page raw --csv > gpurun_out/k1_r02_persistent_raw.csv 2>/dev/null
# launch list of the default bench (all kernels, durations only)
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_default_plain.json 2> gpurun_out/ncu_default_plain.err || echo "plain default failed"
timeout 900 $NCU --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_default.log 2>&1
echo "ncu launches rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/*_raw.csv gpurun_out/r02_launches_default.csv
