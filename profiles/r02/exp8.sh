#!/bin/bash
# r02 experiment 8: gpu test tier (shipped build, then the bounds build), K2 finalize CTA width A/B, default bench line v2
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_exp8_tests.log; cat gpurun_out/r02_exp8_tests.log
bash profiles/r02/bounds_check.sh 2>&1 | tail -14
B="python bench.py --workload batch_bf16 --no-cpu-baseline --no-e2e --no-parity --steps 12 --warmup 3"
for rep in 1 2; do
  CADENCE_FIN_WARPS=32 $B > gpurun_out/k2_fin32_r$rep.json 2>/dev/null || echo "fin32 rc=$?"
  $B > gpurun_out/k2_fin8_r$rep.json 2>/dev/null || echo "fin8 rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k2_fin*.json')):
    d=json.load(open(f)); r=d['roofline']
    print(f.split('/')[-1], 'step', round(d['ms_per_step'],3), 'gemm', round(r['gemm_ms_per_step'],3), 'rest', round(d['ms_per_step']-r['gemm_ms_per_step'],3))
PY
(time python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default_v2.json 2> gpurun_out/r02_bench_default_v2.err); echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default_v2.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('run', d['config'])['single_query_latency_ms_p50'], d['e2e'])
for k,v in d['sub_records'].items(): print(k, v.get('value'), v.get('ms_per_step'), (v.get('roofline') or {}).get('frac'), v.get('invalid'))
PY
