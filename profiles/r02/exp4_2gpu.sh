#!/bin/bash
# r02 experiment 4 (2 GPUs): sharded tests, then the default bench line at N=2 (K1 strong scaling on 1M rows + the
# 100M-row bf16-only batched sub-record, each with in-run parity against independent oracles)
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -15 > gpurun_out/r02_exp4_tests.log
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_2gpu_v1.json 2> gpurun_out/r02_bench_2gpu_v1.err); echo "bench rc=$?"
tail -c 2500 gpurun_out/r02_bench_2gpu_v1.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_2gpu_v1.json'))
    print('K1 N=2', d['value'], d['ms_per_step'], d.get('run', d['config'])['single_query_latency_ms_p50'], d['roofline']['frac'], json.dumps(d['parity'])[:1500])
    b=d['sub_records']['batch_bf16']
    print('K2 N=2', b['value'], b['ms_per_step'], b['roofline']['achieved'], b['roofline']['frac'], b['config']['resident'], json.dumps(b['parity'])[:2500])
    print(b['config'])
except Exception as e:
    print('ERR', e)
PY
cat gpurun_out/r02_exp4_tests.log
