#!/bin/bash
# K2 A/B: dry-run modes and the r02 base library, interleaved twice
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_engine.py -x -q -k "growth_reaches or concurrent_append or tag_filter_beyond or 10m_c_oracle" 2>&1 | tail -40 > gpurun_out/r02_newtests.log
python -m pytest tests/test_gpu_batch_bf16.py -x -q 2>&1 | tail -15 >> gpurun_out/r02_newtests.log
for rep in 1 2; do
  for mode in 0 1 2 3; do
    CADENCE_K2_DRYRUN=$mode python bench.py --workload batch_bf16 --no-cpu-baseline --no-e2e --no-parity --steps 12 --warmup 3 > gpurun_out/k2_ab_new_d${mode}_r${rep}.json 2> gpurun_out/k2_ab_new_d${mode}_r${rep}.err || echo "mode $mode rc=$?"
  done
  CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_r02base.so python bench.py --workload batch_bf16 --no-cpu-baseline --no-e2e --no-parity --steps 12 --warmup 3 > gpurun_out/k2_ab_base_r${rep}.json 2> gpurun_out/k2_ab_base_r${rep}.err || echo "base rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/k2_ab_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['ms_per_step'],3), round(r['gemm_ms_per_step'],3), r['segment_launch_ms_last_step'], d['clocks']['sm_mhz'], d['clocks']['power_w_max'])
    except Exception as e:
        print(f, 'ERR', e)
PY
