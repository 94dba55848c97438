#!/bin/bash
# 8-GPU box: scaling points of the contract line (K1, 1M rows, strong scaling) with both exchange transports,
# single-query latency at 8 GPUs, and BASELINE configs[4] (100M x 1024 bf16 over 8 and 4 GPUs).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
T=${1:-e10}
run() { # N extra-args... ; output name in $OUT
  local n=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n "$@"
}
for n in 8 4 2; do
  CADENCE_EXCHANGE=peer run $n --steps 30 --warmup 5 > gpurun_out/${T}_k1_${n}gpu_peer.json 2> gpurun_out/${T}_k1_${n}gpu_peer.err; echo "k1 $n peer rc=$?"
done
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_k1_1gpu.json 2>/dev/null; echo "k1 1 rc=$?"
CADENCE_EXCHANGE=nccl run 8 --steps 30 --warmup 5 > gpurun_out/${T}_k1_8gpu_nccl.json 2> gpurun_out/${T}_k1_8gpu_nccl.err; echo "k1 8 nccl rc=$?"
for ex in peer nccl; do
  CADENCE_EXCHANGE=$ex run 8 --steps 300 --warmup 20 --queries-per-step 1 > gpurun_out/${T}_k1_8gpu_q1_$ex.json 2> gpurun_out/${T}_k1_8gpu_q1_$ex.err; echo "k1 8 q1 $ex rc=$?"
done
CADENCE_EXCHANGE=peer run 8 --workload batch_bf16 --rows 100000000 --steps 10 --warmup 2 > gpurun_out/${T}_k2_8gpu_100m.json 2> gpurun_out/${T}_k2_8gpu_100m.err; echo "k2 8 rc=$?"
CADENCE_EXCHANGE=peer run 4 --workload batch_bf16 --rows 100000000 --bf16-only --steps 6 --warmup 2 > gpurun_out/${T}_k2_4gpu_100m.json 2> gpurun_out/${T}_k2_4gpu_100m.err; echo "k2 4 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_*.json')):
    try:
        t=open(f).read(); j=json.loads(t[t.index('{"metric"'):]); r=j.get('roofline') or {}
        print(f.split('/')[-1], 'value',round(j['value'],1),'ms/step',round(j.get('ms_per_step',0),4),'e2e',round((j.get('e2e') or {}).get('value',0),1),'frac',round(r.get('frac') or 0,3),'lat',j['config'].get('single_query_latency_ms_p50'), j['config'].get('exchange'), 'recall', j['config'].get('recall_at_50_vs_exact_fp32_lane'))
    except Exception as e: print(f,'ERR',e)
PY
