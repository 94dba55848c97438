#!/bin/bash
# r02 experiment 9: K2 segment growth factor A/B with the round-2 epilogue (4 = default)
cd $GRAFT_REPO_ROOT
B="python bench.py --workload batch_bf16 --no-cpu-baseline --no-e2e --no-parity --steps 12 --warmup 3"
for rep in 1 2; do
  for g in 4 6 8 12; do
    CADENCE_K2_GROWTH=$g $B > gpurun_out/k2_growth${g}_r$rep.json 2>/dev/null || echo "growth $g rc=$?"
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k2_growth*.json')):
    d=json.load(open(f)); r=d['roofline']
    print(f.split('/')[-1], 'step', round(d['ms_per_step'],3), 'gemm', round(r['gemm_ms_per_step'],3), r['segment_launch_ms_last_step'])
PY
