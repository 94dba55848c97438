#!/bin/bash
# r02 experiment 19 (1 GPU): fused finalize+exchange and PDL build -- GPU tier, then single-query latency A/B at N = 1
cd $GRAFT_REPO_ROOT
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02_exp19_tests.log 2>&1
for pdl in 1 0 1 0; do
  CADENCE_PDL=$pdl python profiles/r02/latency/latency_probe.py >> gpurun_out/r02_exp19_latency_1gpu.jsonl 2>> gpurun_out/r02_exp19_latency.err
done
cat gpurun_out/r02_exp19_tests.log
cat gpurun_out/r02_exp19_latency_1gpu.jsonl
tail -5 gpurun_out/r02_exp19_latency.err
