#!/bin/bash
# r02 experiment 35 (1 GPU): output views built AFTER the launches are enqueued (one allocation, pointer arithmetic) --
# GPU tier, then single-query latency against the previous Python side (same library) on one box
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2 3; do
  PROBE_PKG_ROOT=$GRAFT_REPO_ROOT/build/old_pkg python profiles/r02/latency/latency_probe.py 2>> gpurun_out/r02_exp35.err | grep '^{' | sed 's/^{/{"py": "views before the call", /' >> gpurun_out/r02_exp35_latency_1gpu.jsonl
  python profiles/r02/latency/latency_probe.py 2>> gpurun_out/r02_exp35.err | grep '^{' | sed 's/^{/{"py": "views after the call", /' >> gpurun_out/r02_exp35_latency_1gpu.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_exp35_latency_1gpu.jsonl'):
    d=json.loads(l); print(d['py'], 'exact p50 %.4f min %.4f | scan_bf16 p50 %.4f min %.4f' % (d['exact']['p50_ms'], d['exact']['min_ms'], d['scan_bf16']['p50_ms'], d['scan_bf16']['min_ms']))
PY
tail -2 gpurun_out/r02_exp35.err
