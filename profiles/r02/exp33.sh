#!/bin/bash
# r02 experiment 33 (1 GPU): shared passes on the bf16-row scan lane (2 or 4 queries per pass) -- tests, then ms per call
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_engine.py tests/test_gpu_sharded.py -x -q -k "bf16 or scan or shard or rank or hybrid" 2>&1 | tail -4
for sh in "" 0 2 4; do
  if [ -z "$sh" ]; then unset CADENCE_BF16_SHARE; else export CADENCE_BF16_SHARE=$sh; fi
  python profiles/r02/bf16_share_probe.py 2>> gpurun_out/r02_exp33.err >> gpurun_out/r02_exp33_bf16_share.jsonl
done
unset CADENCE_BF16_SHARE
python - <<'PY'
import json
for l in open('gpurun_out/r02_exp33_bf16_share.jsonl'):
    d=json.loads(l)
    print('share', d['share'], {k:{a:round(b,3) for a,b in v.items()} for k,v in d.items() if k not in ('share','rows')})
PY
tail -3 gpurun_out/r02_exp33.err
