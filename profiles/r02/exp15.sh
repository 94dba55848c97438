#!/bin/bash
# r02 experiment 15: per-segment epilogue groups (auto) x cluster form (2 = multicast, 3 = 2-SM MMA), interleaved reps, same box
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_batch_bf16.py -x -q 2>&1 | tail -2
CADENCE_K2_CLUSTER=3 python -m pytest tests/test_gpu_batch_bf16.py -x -q 2>&1 | tail -2
B="python bench.py --workload batch_bf16 --no-cpu-baseline --no-e2e --steps 12 --warmup 3"
for rep in 1 2 3; do
  for c in 2 3; do
    CADENCE_K2_CLUSTER=$c $B > gpurun_out/k2_auto_c${c}_r$rep.json 2>/dev/null || echo "auto c$c rc=$?"
    CADENCE_K2_CLUSTER=$c CADENCE_K2_EPI=1 $B > gpurun_out/k2_epi1_c${c}_r$rep.json 2>/dev/null || echo "epi1 c$c rc=$?"
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k2_auto_c*.json')+glob.glob('gpurun_out/k2_epi1_c*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],3), 'gemm', round(r['gemm_ms_per_step'],3), r['segment_launch_ms_last_step'], 'frac', round(r['frac'],4), 'recall', d['parity']['torch_fp32_matmul_over_fp32_rows']['recall_at_50'], d['parity']['torch_fp32_matmul_over_fp32_rows']['identical_positions'])
    except Exception as e: print(f, 'ERR', e)
PY
