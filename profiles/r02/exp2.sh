#!/bin/bash
# r02 experiment 2: full gpu test tier; K2 limiter hunt (dry-run aids x cluster modes); bf16 scan FHFMA vs widening
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_exp2_tests.log
B="python bench.py --workload batch_bf16 --no-cpu-baseline --no-e2e --no-parity"
for cl in 2 3; do
  for mode in 0 1 4 5; do
    CADENCE_K2_CLUSTER=$cl CADENCE_K2_DRYRUN=$mode $B --steps 10 --warmup 3 > gpurun_out/k2_lim_c${cl}_d${mode}.json 2> gpurun_out/k2_lim_c${cl}_d${mode}.err || echo "c$cl d$mode rc=$?"
  done
done
for rep in 1 2; do
  $B --steps 3 --warmup 3 > gpurun_out/bf16scan_fhfma_r$rep.json 2>/dev/null || echo "fhfma rc=$?"
  CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_bf16widen.so $B --steps 3 --warmup 3 > gpurun_out/bf16scan_widen_r$rep.json 2>/dev/null || echo "widen rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k2_lim_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['ms_per_step'],3), round(r['gemm_ms_per_step'],3), r['segment_launch_ms_last_step'], d['clocks']['sm_mhz'], d['clocks']['power_w_max'])
    except Exception as e:
        print(f, 'ERR', e)
for f in sorted(glob.glob('gpurun_out/bf16scan_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], d['config']['ann_bf16_scan_single_query'], d['config']['exact_fp32_lane_single_query'])
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/r02_exp2_tests.log
