#!/bin/bash
# r02 experiment 18: K2 epilogue warps sleep between polls of the accumulator barrier (power): 0 / 100 / 400 / 1500 ns, interleaved
cd $GRAFT_REPO_ROOT
B="python bench.py --workload batch_bf16 --no-cpu-baseline --no-e2e --steps 12 --warmup 3"
for rep in 1 2 3; do
  for ns in 0 100 400 1500; do
    CADENCE_K2_EPI_SLEEP=$ns $B > gpurun_out/k2_sleep${ns}_r$rep.json 2>/dev/null || echo "sleep $ns rc=$?"
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k2_sleep*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], 'step', round(d['ms_per_step'],3), 'gemm', round(r['gemm_ms_per_step'],3), r['segment_launch_ms_last_step'], 'frac', round(r['frac'],4), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], 'recall', d['parity']['torch_fp32_matmul_over_fp32_rows']['recall_at_50'])
    except Exception as e: print(f, 'ERR', e)
PY
