#!/bin/bash
# r02 experiment 29 (1 GPU): round-end rehearsal of the final build -- smoke(), the gpu test tier, the reference arm, the
# default line -- then ncu --set full of K1 (64 queries / one query / bf16 rows) and of the latency finalize on this build
cd $GRAFT_REPO_ROOT
NCU=/usr/local/cuda/bin/ncu
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
(time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_v7.json 2> gpurun_out/r02_bench_reference_v7.err); echo "reference rc=$?"
(time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_default_v7.json 2> gpurun_out/r02_bench_default_v7.err); echo "bench rc=$?"
python - <<'PY'
import json
r=json.load(open('gpurun_out/r02_bench_reference_v7.json')); d=json.load(open('gpurun_out/r02_bench_default_v7.json'))
print('reference', r['value'], r['ms_per_step'], r['cpu_baseline']['cores'], r['cpu_baseline']['sample'][:80])
print('same config:', r['config'] == d['config'])
print('b200', d['value'], d['e2e']['value'], 'ratio', d['value']/r['value'], 'e2e ratio', d['e2e']['value']/r['value'], d['roofline']['frac'], d['run'], d['clocks'])
for k,v in d['sub_records'].items(): print(k, v.get('value'), v.get('ms_per_step'), (v.get('roofline') or {}).get('frac'), v.get('invalid'), list((v.get('parity') or {}).keys()))
print(d['cpu_baseline'], d['cpu_baseline_hnsw'])
PY
