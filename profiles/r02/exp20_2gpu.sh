#!/bin/bash
# r02 experiment 20 (2 GPUs): fused finalize+exchange / PDL -- sharded tests, single-query latency A/B, finalize phase
# timing (probe build), then the default bench line at N = 2
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(time python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -8) > gpurun_out/r02_exp20_tests.log 2>&1
port=29600
for rep in 1 2; do
for combo in "1 1" "0 1" "1 0" "0 0"; do
  set -- $combo; port=$((port+1))
  CADENCE_PEER_FUSED=$1 CADENCE_PDL=$2 $TR --master-port $port profiles/r02/latency/latency_probe.py >> gpurun_out/r02_exp20_latency_2gpu.jsonl 2>> gpurun_out/r02_exp20_latency.err
done
done
PROBE_ITERS=12 CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_fintiming.so $TR --master-port 29650 profiles/r02/latency/latency_probe.py > gpurun_out/r02_exp20_fintiming_2gpu.log 2>&1
PROBE_ITERS=12 CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_fintiming.so python profiles/r02/latency/latency_probe.py > gpurun_out/r02_exp20_fintiming_1gpu.log 2>&1
(time $TR --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_2gpu_v2.json 2> gpurun_out/r02_bench_2gpu_v2.err); echo "bench rc=$?"
cat gpurun_out/r02_exp20_tests.log
cat gpurun_out/r02_exp20_latency_2gpu.jsonl
tail -3 gpurun_out/r02_exp20_latency.err
grep FIN gpurun_out/r02_exp20_fintiming_2gpu.log | tail -30
grep FIN gpurun_out/r02_exp20_fintiming_1gpu.log | tail -12
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_2gpu_v2.json'))
    print('K1 N=2', d['value'], d['ms_per_step'], d['run'], d['roofline']['frac'])
    b=d['sub_records']['batch_bf16']
    print('K2 N=2', b['value'], b['ms_per_step'], b['roofline']['achieved'], b['roofline']['frac'])
except Exception as e:
    print('ERR', e)
PY
tail -c 600 gpurun_out/r02_bench_2gpu_v2.err
