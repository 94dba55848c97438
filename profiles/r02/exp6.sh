#!/bin/bash
# r02 experiment 6: K2 mainloop probes -- no query-tile reloads (d6), no operand traffic at all (d7) -- next to d5 (no epilogue)
cd $GRAFT_REPO_ROOT
B="python bench.py --workload batch_bf16 --no-cpu-baseline --no-e2e --no-parity"
for rep in 1 2; do
for cl in 2 3; do
  for mode in 5 6 7; do
    CADENCE_K2_CLUSTER=$cl CADENCE_K2_DRYRUN=$mode $B --steps 10 --warmup 3 > gpurun_out/k2_ml_c${cl}_d${mode}_r$rep.json 2> gpurun_out/k2_ml_c${cl}_d${mode}_r$rep.err || echo "c$cl d$mode rc=$?"
  done
done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k2_ml_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['ms_per_step'],3), round(r['gemm_ms_per_step'],3), r['segment_launch_ms_last_step'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))
    except Exception as e:
        print(f, 'ERR', e)
PY
