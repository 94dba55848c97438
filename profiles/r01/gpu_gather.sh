#!/bin/bash
# selective-filter gather path: full GPU test tier, then hybrid + default bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/e11_pytest.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/e11_pytest.log
python bench.py --workload hybrid --steps 40 > gpurun_out/e11_hybrid.json 2> gpurun_out/e11_hybrid.err; echo "hybrid rc=$?"
CADENCE_K1_GATHER=0 python bench.py --workload hybrid --steps 40 > gpurun_out/e11_hybrid_nogather.json 2>/dev/null; echo "hybrid nogather rc=$?"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/e11_default.json 2>/dev/null; echo "default rc=$?"
python - <<PY
import json
for f in ('e11_hybrid','e11_hybrid_nogather'):
    j=json.load(open('gpurun_out/%s.json'%f)); print(f, json.dumps(j['hybrid']))
j=json.load(open('gpurun_out/e11_default.json')); print('default', j['value'], j['roofline']['frac'])
PY
