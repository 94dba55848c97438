#!/bin/bash
# r02 experiment 12: ncu --set full of the EARLY K2 segment launches (segments 2, 3, 4 of a batch), real vs no-append
cd $GRAFT_REPO_ROOT
NCU=/usr/local/cuda/bin/ncu
B="python bench.py --workload batch_bf16 --no-cpu-baseline --no-e2e --no-parity --steps 1 --warmup 2"
for mode in 0 1; do
  CADENCE_K2_DRYRUN=$mode timeout 600 $NCU --set full --clock-control none -k regex:gemm_topk_kernel --launch-skip 13 --launch-count 3 -f -o gpurun_out/k2_r02_early_d$mode $B > gpurun_out/ncu_k2_early_d$mode.log 2>&1
  echo "ncu d$mode rc=$?"
  $NCU -i gpurun_out/k2_r02_early_d$mode.ncu-rep --page raw --csv > gpurun_out/k2_r02_early_d${mode}_raw.csv 2>/dev/null
done
ls -la gpurun_out/k2_r02_early_*
