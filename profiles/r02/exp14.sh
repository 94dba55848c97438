#!/bin/bash
# r02 experiment 14: K2 epilogue warp groups (CADENCE_K2_EPI = 1 | 2 | 4) and first-segment size, same box
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_batch_bf16.py -x -q 2>&1 | tail -3
CADENCE_K2_EPI=4 python -m pytest tests/test_gpu_batch_bf16.py -x -q 2>&1 | tail -2
CADENCE_K2_EPI=1 python -m pytest tests/test_gpu_batch_bf16.py -x -q 2>&1 | tail -2
B="python bench.py --workload batch_bf16 --no-cpu-baseline --no-e2e --steps 12 --warmup 3"
for rep in 1 2; do
  for epi in 1 2 4; do
    CADENCE_K2_EPI=$epi $B > gpurun_out/k2_epi${epi}_r$rep.json 2>/dev/null || echo "epi$epi rc=$?"
  done
  CADENCE_K2_EPI=2 CADENCE_K2_SEG0=8 CADENCE_K2_GROWTH=5 $B > gpurun_out/k2_epi2_s8g5_r$rep.json 2>/dev/null || echo "s8g5 rc=$?"
  CADENCE_K2_EPI=2 CADENCE_K2_GROWTH=6 $B > gpurun_out/k2_epi2_g6_r$rep.json 2>/dev/null || echo "g6 rc=$?"
  CADENCE_K2_EPI=2 CADENCE_K2_CLUSTER=3 $B > gpurun_out/k2_epi2_c3_r$rep.json 2>/dev/null || echo "c3 rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k2_epi[124]_*.json')+glob.glob('gpurun_out/k2_epi2_*_r*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],3), 'gemm', round(r['gemm_ms_per_step'],3), r['segment_launch_ms_last_step'], 'frac', round(r['frac'],4), 'recall', d['parity']['torch_fp32_matmul_over_fp32_rows']['recall_at_50'], d['parity']['torch_fp32_matmul_over_fp32_rows']['identical_positions'])
    except Exception as e: print(f, 'ERR', e)
PY
