#!/bin/bash
# r02 experiment 25 (2 GPUs): fused exchange + st.async latency finalize + merge by binary search -- sharded tests, then
# single-query latency interleaved with the library of commit ed8c491 on the same box, and the finalize's phase timing
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(time python -m pytest tests/test_gpu_sharded.py tests/test_gpu_engine.py -x -q -k "shard or merge or rank" 2>&1 | tail -8) > gpurun_out/r02_exp25_tests.log 2>&1
cat gpurun_out/r02_exp25_tests.log
port=29700
for i in 1 2 3; do
  port=$((port+1))
  CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_base_ed8c491.so $TR --master-port $port profiles/r02/latency/latency_probe.py 2>> gpurun_out/r02_exp25_latency.err | grep '^{' | sed 's/^{/{"lib": "ed8c491", /' >> gpurun_out/r02_exp25_latency_2gpu.jsonl
  port=$((port+1))
  $TR --master-port $port profiles/r02/latency/latency_probe.py 2>> gpurun_out/r02_exp25_latency.err | grep '^{' | sed 's/^{/{"lib": "new", /' >> gpurun_out/r02_exp25_latency_2gpu.jsonl
done
port=$((port+1))
CADENCE_PEER_FUSED=0 $TR --master-port $port profiles/r02/latency/latency_probe.py 2>> gpurun_out/r02_exp25_latency.err | grep '^{' | sed 's/^{/{"lib": "new, separate K4p launch", /' >> gpurun_out/r02_exp25_latency_2gpu.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r02_exp25_latency_2gpu.jsonl'):
    d=json.loads(l); print(d['lib'], 'exact p50 %.4f min %.4f | scan_bf16 p50 %.4f min %.4f | batch64 %.2f %.2f' % (d['exact']['p50_ms'], d['exact']['min_ms'], d['scan_bf16']['p50_ms'], d['scan_bf16']['min_ms'], d['exact']['batch64_ms'], d['scan_bf16']['batch64_ms']))
PY
PROBE_ITERS=8 CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_fintiming.so $TR --master-port 29750 profiles/r02/latency/latency_probe.py > gpurun_out/r02_exp25_fintiming_2gpu.log 2>&1
grep -o "rank [0-9]* (count.*" gpurun_out/r02_exp25_fintiming_2gpu.log | tail -12
