#!/bin/bash
# r02 experiment 7 (N GPUs): the default bench line at N = $1 (K1 strong scaling on 1M rows + the 100M-row batched sub-record)
N=${1:-8}
cd $GRAFT_REPO_ROOT
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu_v1.json 2> gpurun_out/r02_bench_${N}gpu_v1.err); echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench_${N}gpu_v1.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_${N}gpu_v1.json'))
    print('K1', d['n_gpus'], d['value'], d['ms_per_step'], 'lat', d.get('run', d['config'])['single_query_latency_ms_p50'], 'frac', d['roofline']['frac'], 'avg_launch_ms', d['roofline']['avg_launch_ms'], d['clocks'])
    print(' parity', {k:(v.get('identical_positions', v.get('identical'))) for k,v in d['parity'].items()})
    print(' shared', d['exact_batch_shared_reads'])
    b=d['sub_records']['batch_bf16']
    print('K2', b['value'], b['ms_per_step'], b['roofline']['achieved'], b['roofline']['frac'], b['config']['resident'], b['clocks'])
    print(' parity', json.dumps(b['parity'])[:1800])
    print(' cfg', {k:v for k,v in b['config'].items() if k!='workload'})
except Exception as e:
    print('ERR', e)
PY
