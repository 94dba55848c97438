#!/bin/bash
# r02 experiment 26 (1 GPU): tech lane by compaction + distinct merges, branch-free RRF -- GPU tier, default bench, launch list
cd $GRAFT_REPO_ROOT
NCU=/usr/local/cuda/bin/ncu
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/r02_exp26_tests.log 2>&1
cat gpurun_out/r02_exp26_tests.log
(time python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_v5.json 2> gpurun_out/r02_bench_1gpu_v5.err); echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_1gpu_v5.json'))
    print('K1', d['value'], d['ms_per_step'], d['run'], d['roofline']['frac'], 'e2e', d['e2e']['value'])
    b=d['sub_records']['batch_bf16']
    print('K2', b['value'], b['ms_per_step'], b['roofline']['achieved'], b['roofline']['frac'], json.dumps(b.get('run'))[:600])
    h=d['sub_records']['hybrid']
    print('hybrid', h['value'], json.dumps(h['hybrid'])[:1800])
except Exception as e:
    print('ERR', e)
PY
tail -c 300 gpurun_out/r02_bench_1gpu_v5.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_default_plain.json 2> gpurun_out/ncu_default_plain.err || echo "plain default failed"
timeout 900 $NCU --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r02_launches_default_v2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_default.log 2>&1
echo "ncu launches rc=$?"
