#!/bin/bash
# 1-GPU evidence pass of the current tree: full GPU test tier, smoke, every bench workload, reference arm.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
T=${1:-e9}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "default rc=$?"
python bench.py --queries-per-step 1 --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/${T}_bench_k1_q1.json 2>/dev/null; echo "q1 rc=$?"
python bench.py --rows 10000000 --queries-per-step 1 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_k1_10m_q1.json 2>/dev/null; echo "10m q1 rc=$?"
python bench.py --workload batch_bf16 --steps 20 --warmup 3 > gpurun_out/${T}_bench_k2.json 2> gpurun_out/${T}_bench_k2.err; echo "k2 rc=$?"
for nq in 16 128 256 512; do
  python bench.py --workload batch_bf16 --batch-queries $nq --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_k2_nq$nq.json 2>/dev/null; echo "k2 nq=$nq rc=$?"
done
python bench.py --workload hybrid --steps 40 > gpurun_out/${T}_bench_hybrid.json 2> gpurun_out/${T}_bench_hybrid.err; echo "hybrid rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>/dev/null; echo "reference rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_bench_*.json')):
    try:
        j=json.load(open(f)); r=j.get('roofline') or {}
        print(f.split('/')[-1], 'value',round(j['value'],1),'ms/step',round(j.get('ms_per_step',0),4),'e2e',(j.get('e2e') or {}).get('value'),'frac',r.get('frac'),'achieved',r.get('achieved'), 'stream', r.get('corpus_stream_gbs'), 'recall', j['config'].get('recall_at_50_vs_exact_fp32_lane'))
    except Exception as e: print(f,'ERR',e)
PY
