#!/bin/bash
# r02 experiment 11: what the driver runs at round end -- smoke(), the gpu test tier, the reference arm, the default line
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
(time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_v3.json 2> gpurun_out/r02_bench_reference_v3.err); echo "reference rc=$?"
(time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_default_v3.json 2> gpurun_out/r02_bench_default_v3.err); echo "bench rc=$?"
python - <<'PY'
import json
r=json.load(open('gpurun_out/r02_bench_reference_v3.json')); d=json.load(open('gpurun_out/r02_bench_default_v3.json'))
print('reference', r['value'], r['ms_per_step'], r['cpu_baseline']['cores'], r['cpu_baseline']['sample'][:80])
print('same config:', r['config'] == d['config'])
print('b200', d['value'], d['e2e']['value'], 'ratio', d['value']/r['value'], 'e2e ratio', d['e2e']['value']/r['value'], d['roofline']['frac'], d['run'])
for k,v in d['sub_records'].items(): print(k, v.get('value'), v.get('ms_per_step'), (v.get('roofline') or {}).get('frac'), v.get('invalid'), list((v.get('parity') or {}).keys()))
print(d['cpu_baseline'], d['cpu_baseline_hnsw'])
PY
