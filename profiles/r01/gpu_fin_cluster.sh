#!/bin/bash
# cluster finalize: tests, then single-query A/B (CADENCE_FIN_CLUSTER=0 = one CTA per query)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py -m gpu -x -q > gpurun_out/e8_pytest.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/e8_pytest.log
for c in 1 0; do
  CADENCE_FIN_CLUSTER=$c python bench.py --steps 300 --warmup 20 --queries-per-step 1 --no-cpu-baseline > gpurun_out/e8_q1_c$c.json 2>/dev/null
  python -c "
import json; j=json.load(open('gpurun_out/e8_q1_c$c.json')); print('cluster=$c ms/step', round(j['ms_per_step'],4), 'k1', round(j['roofline']['avg_launch_ms'],4), 'lat', round(j['config']['single_query_latency_ms_p50'],4), 'e2e q/s', round(j['e2e']['value'],1))"
done
CADENCE_FIN_CLUSTER=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:scan_finalize -c 30 --csv --log-file gpurun_out/e8_fin_cluster.csv \
     python bench.py --steps 10 --warmup 3 --queries-per-step 1 --no-cpu-baseline --no-e2e > gpurun_out/e8_fin_cluster.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/e8_fin_cluster.csv')) if len(r)>10]
i=rows[0].index('Metric Value'); v=sorted(float(r[i]) for r in rows[1:])
print('cluster finalize ns: min',v[0],'median',v[len(v)//2],'max',v[-1],'n',len(v), rows[1][rows[0].index('Kernel Name')][:60])
PY
