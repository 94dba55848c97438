#!/bin/bash
# r02 experiment 10 (N GPUs): pipelined sharded step (finalize + exchange of chunk c beside the scan of chunk c+1) A/B
N=${1:-2}
cd $GRAFT_REPO_ROOT
if [ "$N" = "2" ]; then python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -5; fi
for rep in ${REPS:-1 2}; do
for pipe in 0 1; do
  CADENCE_SHARD_PIPELINE=$pipe python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$pipe bench.py --gpus $N --steps 20 --warmup 5 --no-sub-records > gpurun_out/r02_pipe${pipe}_${N}gpu_r$rep.json 2> gpurun_out/r02_pipe${pipe}_${N}gpu_r$rep.err || { echo "pipe=$pipe rc=$?"; tail -5 gpurun_out/r02_pipe${pipe}_${N}gpu_r$rep.err; }
done
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_pipe*_${N}gpu_r*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], 'q/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'k1 ms/step', round(r.get('k1_ms_per_step', r['avg_launch_ms']),4), 'frac', round(r['frac'],4), 'lat', round(d.get('run', d['config'])['single_query_latency_ms_p50'],4), 'parity', [v.get('identical_positions', v.get('identical')) for v in d['parity'].values()])
    except Exception as e:
        print(f, 'ERR', e)
PY
