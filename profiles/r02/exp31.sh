#!/bin/bash
# r02 experiment 31 (1 GPU): latency finalize capped at 112 registers (U = 8, dim <= 1024) so that it is co-resident with the
# scan under programmatic dependent launch -- interleaved with the library of commit 8ebd0a5 (222 registers) on one box
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_engine.py -x -q -k "finalize or exact or scan" 2>&1 | tail -3
for i in 1 2 3; do
  CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_head_8ebd0a5.so python profiles/r02/latency/latency_probe.py 2>> gpurun_out/r02_exp31.err | grep '^{' | sed 's/^{/{"lib": "8ebd0a5 (222 regs)", /' >> gpurun_out/r02_exp31_latency_1gpu.jsonl
  python profiles/r02/latency/latency_probe.py 2>> gpurun_out/r02_exp31.err | grep '^{' | sed 's/^{/{"lib": "112 regs", /' >> gpurun_out/r02_exp31_latency_1gpu.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_exp31_latency_1gpu.jsonl'):
    d=json.loads(l); print(d['lib'], 'exact p50 %.4f min %.4f | scan_bf16 p50 %.4f min %.4f | batch64 %.2f %.2f' % (d['exact']['p50_ms'], d['exact']['min_ms'], d['scan_bf16']['p50_ms'], d['scan_bf16']['min_ms'], d['exact']['batch64_ms'], d['scan_bf16']['batch64_ms']))
PY
