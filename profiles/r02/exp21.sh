#!/bin/bash
# r02 experiment 21 (1 GPU): integer order keys in every ranking loop -- GPU tier, latency, finalize phase timing, default bench
cd $GRAFT_REPO_ROOT
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02_exp21_tests.log 2>&1
for i in 1 2; do python profiles/r02/latency/latency_probe.py >> gpurun_out/r02_exp21_latency_1gpu.jsonl 2>> gpurun_out/r02_exp21_latency.err; done
PROBE_ITERS=6 CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_fintiming.so python profiles/r02/latency/latency_probe.py > gpurun_out/r02_exp21_fintiming_1gpu.log 2>&1
(time python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_v4.json 2> gpurun_out/r02_bench_1gpu_v4.err); echo "bench rc=$?"
cat gpurun_out/r02_exp21_tests.log
cat gpurun_out/r02_exp21_latency_1gpu.jsonl
grep FIN gpurun_out/r02_exp21_fintiming_1gpu.log | tail -8
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_1gpu_v4.json'))
    print('K1', d['value'], d['ms_per_step'], d['run'], d['roofline']['frac'], 'e2e', d['e2e']['value'])
    b=d['sub_records']['batch_bf16']
    print('K2', b['value'], b['ms_per_step'], b['roofline']['achieved'], b['roofline']['frac'], {k:v for k,v in b.get('run',{}).items() if 'ms' in k})
    h=d['sub_records']['hybrid']
    print('hybrid', json.dumps(h)[:1500])
except Exception as e:
    print('ERR', e)
PY
tail -c 400 gpurun_out/r02_bench_1gpu_v4.err
