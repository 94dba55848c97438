#!/bin/bash
# 2-GPU pass: peer-memory exchange (K4p) vs NCCL all-gather + K4, tests first.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/e4_pytest_sharded.log 2>&1; echo "sharded rc=$?"; tail -15 gpurun_out/e4_pytest_sharded.log
timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "fused_hybrid" > gpurun_out/e4_pytest_fused.log 2>&1; echo "fused rc=$?"; tail -3 gpurun_out/e4_pytest_fused.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for ex in peer nccl; do
  CADENCE_EXCHANGE=$ex timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/e4_k1_2gpu_$ex.json 2> gpurun_out/e4_k1_2gpu_$ex.err; echo "k1 $ex rc=$?"
  CADENCE_EXCHANGE=$ex timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 --queries-per-step 1 > gpurun_out/e4_k1_2gpu_q1_$ex.json 2> gpurun_out/e4_k1_2gpu_q1_$ex.err; echo "k1 q1 $ex rc=$?"
done
grep -h -o '"value": [0-9.]*\|single_query_latency_ms_p50": [0-9.]*\|"exchange": "[a-z]*"' gpurun_out/e4_k1_2gpu_*.json
