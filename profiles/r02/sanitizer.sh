#!/bin/bash
# compute-sanitizer over the C-ABI driver (tools/sanitize_driver.cc): memcheck, racecheck, synccheck on small shapes of
# every kernel family -- K1 one-scan-per-query (persistent over queries), shared reads (3 in registers, deep 16), gather
# launch, bf16 scan, K2 with cluster 1 / 2 / 2-SM / 4 (CADENCE_K2_CLUSTER) incl. the device-side overflow re-run, both
# finalize kernels, K4, K4p (one rank), K5, K6, tech lane, fused hybrid call.
# Usage (GPU box): bash profiles/r02/sanitizer.sh   -> gpurun_out/sanitizer_*.log
cd ${GRAFT_REPO_ROOT:-.}
out=gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
run() {   # name, tool, extra env, cases...
  name=$1; tool=$2; envs=$3; shift 3
  echo "== $name ($tool) $envs $*" 
  env $envs timeout 900 $SAN --tool $tool --print-limit 30 --error-exitcode 9 ./build/sanitize_driver "$@" > $out/sanitizer_${name}_${tool}.log 2>&1
  rc=$?
  echo "rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|all requested cases passed' $out/sanitizer_${name}_${tool}.log | tr '\n' ' ')"
}
for tool in memcheck synccheck racecheck; do
  run scan $tool "X=1" k1 k1_shared k1_gather bf16_scan filter
  run lanes $tool "X=1" rrf tech hybrid merge
  for cl in 1 2 3 4; do
    run k2c$cl $tool "CADENCE_K2_CLUSTER=$cl" k2 filter
  done
done
