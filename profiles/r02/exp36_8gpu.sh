#!/bin/bash
# r02 experiment 36 (8 GPUs): single-query latency of the last build, old Python side (views before the call) beside it
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
port=29900
for side in old new old new; do
  port=$((port+1))
  if [ $side = old ]; then export PROBE_PKG_ROOT=$GRAFT_REPO_ROOT/build/old_pkg; else unset PROBE_PKG_ROOT; fi
  timeout 300 $TR --master-port $port profiles/r02/latency/latency_probe.py 2>> gpurun_out/r02_exp36.err | grep '^{' | sed "s/^{/{\"py\": \"$side\", /" >> gpurun_out/r02_exp36_latency_8gpu.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_exp36_latency_8gpu.jsonl'):
    d=json.loads(l); print(d['py'], 'exact p50 %.4f min %.4f | scan_bf16 p50 %.4f min %.4f | batch64 %.3f %.3f' % (d['exact']['p50_ms'], d['exact']['min_ms'], d['scan_bf16']['p50_ms'], d['scan_bf16']['min_ms'], d['exact']['batch64_ms'], d['scan_bf16']['batch64_ms']))
PY
