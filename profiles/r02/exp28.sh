#!/bin/bash
# r02 experiment 28 (1 GPU): cluster form of the tech lane for filtered requests, randomized K4 test -- GPU tier + default bench
cd $GRAFT_REPO_ROOT
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/r02_exp28_tests.log 2>&1
cat gpurun_out/r02_exp28_tests.log
for tc in 1 0; do
CADENCE_TECH_CLUSTER=$tc python bench.py --workload hybrid --steps 20 --warmup 5 > gpurun_out/r02_bench_hybrid_tc$tc.json 2> gpurun_out/r02_bench_hybrid_tc$tc.err; echo "bench rc=$?"
python - <<PY
import json
try:
    h=json.load(open('gpurun_out/r02_bench_hybrid_tc$tc.json'))
    h=h.get('sub_records',{}).get('hybrid',h)
    print('tech cluster $tc: hybrid', h['value'], json.dumps(h['hybrid'])[:900])
except Exception as e:
    print('ERR', e)
PY
done
tail -c 300 gpurun_out/r02_bench_hybrid_tc1.err
