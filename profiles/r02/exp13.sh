#!/bin/bash
# r02 experiment 13: K2 epilogue v3 (one atomic per flush, per-tile flush, float-first admission test) vs v2, same box
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_batch_bf16.py -x -q 2>&1 | tail -3
B="python bench.py --workload batch_bf16 --no-cpu-baseline --no-e2e --steps 12 --warmup 3"
for rep in 1 2 3; do
  CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_k2v2.so $B > gpurun_out/k2_epi_v2_r$rep.json 2>/dev/null || echo "v2 rc=$?"
  $B > gpurun_out/k2_epi_v3_r$rep.json 2>/dev/null || echo "v3 rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k2_epi_v*.json')):
    d=json.load(open(f)); r=d['roofline']
    print(f.split('/')[-1], 'step', round(d['ms_per_step'],3), 'gemm', round(r['gemm_ms_per_step'],3), r['segment_launch_ms_last_step'], 'frac', round(r['frac'],4), 'recall', d['parity']['torch_fp32_matmul_over_fp32_rows']['recall_at_50'], d['parity']['torch_fp32_matmul_over_fp32_rows']['identical_positions'])
PY
