#!/bin/bash
# r02 experiment 24 (1 GPU): latency finalize with st.async + mbarrier exchanges (no cluster barrier behind the dependency
# wait): phase timing, GPU tier, then single-query latency interleaved with the library of commit ed8c491 on the SAME box
cd $GRAFT_REPO_ROOT
PROBE_ITERS=6 CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_fintiming.so python profiles/r02/latency/latency_probe.py > gpurun_out/r02_exp24_fintiming_1gpu.log 2>&1
grep FIN gpurun_out/r02_exp24_fintiming_1gpu.log | tail -6
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02_exp24_tests.log 2>&1
cat gpurun_out/r02_exp24_tests.log
for i in 1 2 3; do
  CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_base_ed8c491.so python profiles/r02/latency/latency_probe.py | sed 's/^{/{"lib": "ed8c491", /' >> gpurun_out/r02_exp24_latency_1gpu.jsonl 2>> gpurun_out/r02_exp24_latency.err
  python profiles/r02/latency/latency_probe.py | sed 's/^{/{"lib": "new", /' >> gpurun_out/r02_exp24_latency_1gpu.jsonl 2>> gpurun_out/r02_exp24_latency.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_exp24_latency_1gpu.jsonl'):
    d=json.loads(l); print(d['lib'], 'exact p50 %.4f min %.4f | scan_bf16 p50 %.4f min %.4f | batch64 %.2f %.2f' % (d['exact']['p50_ms'], d['exact']['min_ms'], d['scan_bf16']['p50_ms'], d['scan_bf16']['min_ms'], d['exact']['batch64_ms'], d['scan_bf16']['batch64_ms']))
PY
