#!/bin/bash
# r02 experiment 30 (4 GPUs): sharded tests (extended cases) and the default bench line of the final build at N = 2 and N = 4
cd $GRAFT_REPO_ROOT
(time python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -5) > gpurun_out/r02_exp30_tests.log 2>&1
cat gpurun_out/r02_exp30_tests.log
for n in 2 4; do
(time CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((n-1))) python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_${n}gpu_v3.json 2> gpurun_out/r02_bench_${n}gpu_v3.err); echo "bench N=$n rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench_${n}gpu_v3.json'))
    print('K1 N=$n', d['value'], d['ms_per_step'], d['run'], d['roofline']['frac'], d['roofline'].get('k1_ms_per_step'), 'e2e', d['e2e']['value'], list(d['parity'].keys()))
    b=d['sub_records']['batch_bf16']
    print('K2 N=$n', b['value'], b['ms_per_step'], b['roofline']['achieved'], b['roofline']['frac'], b['config'].get('resident'), list(b['parity'].keys()))
except Exception as e:
    print('ERR', e)
PY
done
