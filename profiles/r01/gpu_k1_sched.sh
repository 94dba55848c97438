#!/bin/bash
# K1 tile scheduling A/B (1 GPU): static split vs work stealing, single-query and 64-query steps.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "exact_scan or full_size or sharded_equals" > gpurun_out/e5_pytest_k1.log 2>&1; echo "k1 tests rc=$?"; tail -4 gpurun_out/e5_pytest_k1.log
for sched in dynamic static; do
  CADENCE_K1_SCHED=$sched python bench.py --steps 200 --warmup 20 --queries-per-step 1 --no-cpu-baseline > gpurun_out/e5_k1_q1_$sched.json 2>gpurun_out/e5_k1_q1_$sched.err; echo "q1 $sched rc=$?"
  CADENCE_K1_SCHED=$sched python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/e5_k1_q64_$sched.json 2>gpurun_out/e5_k1_q64_$sched.err; echo "q64 $sched rc=$?"
  CADENCE_K1_SCHED=$sched python bench.py --rows 10000000 --steps 20 --warmup 3 --queries-per-step 1 --no-cpu-baseline > gpurun_out/e5_k1_10m_q1_$sched.json 2>gpurun_out/e5_k1_10m_q1_$sched.err; echo "10m q1 $sched rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/e5_k1_*.json')):
    try:
        j=json.load(open(f)); r=j['roofline']
        print(f.split('/')[-1], 'q/s',round(j['value'],1),'ms/step',round(j['ms_per_step'],4),'k1_ms',round(r['avg_launch_ms'],4),'GB/s',round(r['achieved']),'lat',round(j['config']['single_query_latency_ms_p50'],4),'e2e',round(j['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
