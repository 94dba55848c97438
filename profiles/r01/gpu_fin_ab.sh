#!/bin/bash
# finalize kernel: load-batch size A/B (ncu launch durations; cold-cache, comparable with each other only)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "exact_scan or full_size" > gpurun_out/e6_pytest.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/e6_pytest.log
for rb in 1 2 4 8; do
  CADENCE_FIN_RB=$rb ncu --metrics gpu__time_duration.sum --clock-control none -k regex:scan_finalize -c 30 --csv --log-file gpurun_out/e6_fin_rb$rb.csv \
     python bench.py --steps 10 --warmup 3 --queries-per-step 1 --no-cpu-baseline --no-e2e > gpurun_out/e6_fin_rb$rb.log 2>&1
  echo "rb=$rb rc=$?"
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/e6_fin_rb$rb.csv')) if len(r)>10]
i=rows[0].index('Metric Value'); v=sorted(float(r[i]) for r in rows[1:])
print('rb=$rb finalize ns: min',v[0],'median',v[len(v)//2],'max',v[-1],'n',len(v))
PY
done
for rb in 1 2 4 8; do
  CADENCE_FIN_RB=$rb python bench.py --steps 300 --warmup 20 --queries-per-step 1 --no-cpu-baseline > gpurun_out/e6_q1_rb$rb.json 2>/dev/null
  python -c "
import json; j=json.load(open('gpurun_out/e6_q1_rb$rb.json')); print('rb=$rb ms/step', round(j['ms_per_step'],4), 'k1', round(j['roofline']['avg_launch_ms'],4), 'lat', round(j['config']['single_query_latency_ms_p50'],4))"
done
