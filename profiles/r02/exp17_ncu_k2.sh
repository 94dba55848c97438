#!/bin/bash
# r02 experiment 17: current K2 build (2-SM MMA form, two epilogue groups in the small segments): launch list of one
# batch_bf16 step (every kernel, gpu__time_duration) and ncu --set full of the last-segment launch
cd $GRAFT_REPO_ROOT
NCU=/usr/local/cuda/bin/ncu
B="python bench.py --workload batch_bf16 --no-cpu-baseline --no-e2e --no-parity --steps 1 --warmup 2"
$B > gpurun_out/ncu_k2v4_plain.json 2> gpurun_out/ncu_k2v4_plain.err || { echo "plain K2 run failed"; exit 1; }
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_batch_bf16.csv $B > gpurun_out/ncu_k2v4_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:gemm_topk_kernel --launch-skip 17 --launch-count 1 -f -o gpurun_out/k2_r02_v4_last_segment $B > gpurun_out/ncu_k2v4.log 2>&1
echo "ncu k2 rc=$?"
$NCU -i gpurun_out/k2_r02_v4_last_segment.ncu-rep --page raw --csv > gpurun_out/k2_r02_v4_last_segment_raw.csv 2>/dev/null
ls -la gpurun_out/k2_r02_v4* gpurun_out/r02_launches_batch_bf16.csv
