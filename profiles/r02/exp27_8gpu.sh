#!/bin/bash
# r02 experiment 27 (8 GPUs): single-query latency against the library of commit ed8c491 on the same box, then the default
# bench line at N = 8 (1 M rows sharded; sub-record: 100 M rows bf16)
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
port=29800
for lib in base new new base; do
  port=$((port+1))
  if [ $lib = base ]; then export CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_base_ed8c491.so; tag=ed8c491; else unset CADENCE_DENSE_LIB; tag=new; fi
  timeout 300 $TR --master-port $port profiles/r02/latency/latency_probe.py 2>> gpurun_out/r02_exp27_latency.err | grep '^{' | sed "s/^{/{\"lib\": \"$tag\", /" >> gpurun_out/r02_exp27_latency_8gpu.jsonl
done
unset CADENCE_DENSE_LIB
python - <<'PY'
import json
for l in open('gpurun_out/r02_exp27_latency_8gpu.jsonl'):
    d=json.loads(l); print(d['lib'], 'exact p50 %.4f min %.4f | scan_bf16 p50 %.4f min %.4f | batch64 %.3f %.3f' % (d['exact']['p50_ms'], d['exact']['min_ms'], d['scan_bf16']['p50_ms'], d['scan_bf16']['min_ms'], d['exact']['batch64_ms'], d['scan_bf16']['batch64_ms']))
PY
(time timeout 600 $TR --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu_v2.json 2> gpurun_out/r02_bench_8gpu_v2.err); echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_8gpu_v2.json'))
    print('K1 N=8', d['value'], d['ms_per_step'], d['run'], d['roofline']['frac'], d['roofline'].get('k1_ms_per_step'), 'e2e', d['e2e']['value'])
    print(json.dumps(d['parity'])[:800])
    b=d['sub_records']['batch_bf16']
    print('K2 N=8', b['value'], b['ms_per_step'], b['roofline']['achieved'], b['roofline']['frac'], json.dumps(b['parity'])[:600])
except Exception as e:
    print('ERR', e)
PY
tail -c 400 gpurun_out/r02_bench_8gpu_v2.err
