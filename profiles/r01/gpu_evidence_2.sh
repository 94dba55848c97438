#!/bin/bash
# Round-1 evidence pass #2 (run under gpurun, one B200).  Every ncu command runs only after the
# same program exited 0 without ncu; numbers printed under ncu are never quoted as bench values.
set -x
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-/root/repo}"

# 1. contract line (configs[1]) with the time-bounded CPU baseline
python bench.py --steps 20 --warmup 5 > gpurun_out/e2_bench_default.json 2> gpurun_out/e2_bench_default.err
echo "default rc=$?"

# 2. north-star headline: the same kernel over 10M x 1024 fp32 (40.96 GB per query)
python bench.py --rows 10000000 --queries-per-step 8 --steps 6 --warmup 3 --no-cpu-baseline \
    > gpurun_out/e2_bench_k1_10m.json 2> gpurun_out/e2_bench_k1_10m.err
echo "k1_10m rc=$?"

# 3. K2 (configs[2]) plain run, then its launch list, then dram bytes of every gemm launch of one
#    step, then one --set full capture of the largest segment launch
python bench.py --workload batch_bf16 --steps 10 --warmup 3 > gpurun_out/e2_bench_k2.json 2> gpurun_out/e2_bench_k2.err
rc=$?; echo "k2 rc=$rc"
if [ $rc -eq 0 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/e2_k2_launches.csv \
      python bench.py --workload batch_bf16 --steps 2 --warmup 1 --no-e2e > gpurun_out/e2_k2_launches.log 2>&1
  echo "k2 launch list rc=$?"
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
      --clock-control none -k regex:gemm_topk --csv --log-file gpurun_out/e2_k2_dram_per_launch.csv \
      python bench.py --workload batch_bf16 --steps 1 --warmup 1 --no-e2e > gpurun_out/e2_k2_dram.log 2>&1
  echo "k2 dram rc=$?"
fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,temperature.gpu --format=csv > gpurun_out/e2_smi.csv

# 4. A/B: tile-major walk of the large segments (CADENCE_K2_ORDER=1): plain bench, then dram per launch
CADENCE_K2_ORDER=1 python bench.py --workload batch_bf16 --steps 10 --warmup 3 > gpurun_out/e2_bench_k2_order1.json 2> gpurun_out/e2_bench_k2_order1.err
rc=$?; echo "k2 order1 rc=$rc"
if [ $rc -eq 0 ]; then
  CADENCE_K2_ORDER=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
      --clock-control none -k regex:gemm_topk --csv --log-file gpurun_out/e2_k2_dram_per_launch_order1.csv \
      python bench.py --workload batch_bf16 --steps 1 --warmup 1 --no-e2e > gpurun_out/e2_k2_dram_order1.log 2>&1
  echo "k2 order1 dram rc=$?"
fi
# 5. second plain K2 run of the default order (run-to-run spread under the power cap)
python bench.py --workload batch_bf16 --steps 10 --warmup 3 > gpurun_out/e2_bench_k2_b.json 2> gpurun_out/e2_bench_k2_b.err
tail -c 600 gpurun_out/e2_bench_k2.json gpurun_out/e2_bench_k2_order1.json gpurun_out/e2_bench_k2_b.json
