#!/bin/bash
# r02 experiment 3: gpu test tier; K2 corpus-tile multicast width (cluster 2 / 2-SM / 4) with and without the epilogue;
# bf16 scan FHFMA vs widening (both with the joint 4-row reduction); sanitizer driver as a plain run
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_exp3_tests.log
./build/sanitize_driver > gpurun_out/r02_sanitize_driver_plain.log 2>&1; echo "driver rc=$?" >> gpurun_out/r02_sanitize_driver_plain.log
B="python bench.py --workload batch_bf16 --no-cpu-baseline --no-e2e --no-parity"
for rep in 1 2; do
for cl in 2 3 4; do
  for mode in 0 5; do
    CADENCE_K2_VERBOSE=1 CADENCE_K2_CLUSTER=$cl CADENCE_K2_DRYRUN=$mode $B --steps 10 --warmup 3 > gpurun_out/k2_cl_c${cl}_d${mode}_r$rep.json 2> gpurun_out/k2_cl_c${cl}_d${mode}_r$rep.err || echo "c$cl d$mode rc=$?"
  done
done
done
for rep in 1 2; do
  $B --steps 3 --warmup 3 > gpurun_out/bf16scan2_fhfma_r$rep.json 2>/dev/null || echo "fhfma rc=$?"
  CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_bf16widen.so $B --steps 3 --warmup 3 > gpurun_out/bf16scan2_widen_r$rep.json 2>/dev/null || echo "widen rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k2_cl_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['ms_per_step'],3), round(r['gemm_ms_per_step'],3), r['segment_launch_ms_last_step'], d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
for f in sorted(glob.glob('gpurun_out/bf16scan2_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], d['config']['ann_bf16_scan_single_query'], d['config']['exact_fp32_lane_single_query'])
    except Exception as e:
        print(f, 'ERR', e)
PY
grep -h "co-resident" gpurun_out/k2_cl_c4_d0_r1.err | head -2
cat gpurun_out/r02_exp3_tests.log; tail -5 gpurun_out/r02_sanitize_driver_plain.log
