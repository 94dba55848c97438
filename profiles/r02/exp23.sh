#!/bin/bash
# r02 experiment 23 (1 GPU): latency finalize with push-only cluster exchanges, one-batch re-score, branch-free ranking
cd $GRAFT_REPO_ROOT
PROBE_ITERS=6 CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_fintiming.so python profiles/r02/latency/latency_probe.py > gpurun_out/r02_exp23_fintiming_1gpu.log 2>&1
grep FIN gpurun_out/r02_exp23_fintiming_1gpu.log | tail -6
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02_exp23_tests.log 2>&1
cat gpurun_out/r02_exp23_tests.log
for i in 1 2; do python profiles/r02/latency/latency_probe.py >> gpurun_out/r02_exp23_latency_1gpu.jsonl 2>> gpurun_out/r02_exp23_latency.err; done
cat gpurun_out/r02_exp23_latency_1gpu.jsonl
