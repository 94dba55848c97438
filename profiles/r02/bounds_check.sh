#!/bin/bash
# compute-sanitizer is closed on this pool ("runs under it have left GPUs needing a reset"), so the evidence asked for is
# produced with the library's own checks instead: a -DCDR_DEBUG_BOUNDS build (device-side asserts on stage / list /
# staging-buffer / slot indices and pipeline invariants: CDR_DEV_ASSERT in csrc/) runs (a) the C-ABI driver over every
# kernel family, with K2 in its four cluster forms, and (b) the whole GPU test tier.  A failed assert traps: the launch
# errors out, the C ABI returns CDR_ERR_CUDA and the run fails.
cd ${GRAFT_REPO_ROOT:-.}
out=gpurun_out
for cl in 1 2 3 4; do
  echo "== driver, bounds build, CADENCE_K2_CLUSTER=$cl"
  CADENCE_K2_CLUSTER=$cl LD_LIBRARY_PATH=$PWD/build/ab/bounds timeout 600 ./build/sanitize_driver > $out/bounds_driver_k2c$cl.log 2>&1
  echo "rc=$? $(tail -1 $out/bounds_driver_k2c$cl.log)"
done
LD_LIBRARY_PATH=$PWD/build/ab/bounds ldd ./build/sanitize_driver | grep cadence
echo "== gpu test tier, bounds build"
CADENCE_DENSE_LIB=$PWD/build/ab/bounds/libcadence_dense.so timeout 1200 python -m pytest tests -m gpu -x -q > $out/bounds_pytest_gpu.log 2>&1
echo "rc=$? $(tail -1 $out/bounds_pytest_gpu.log)"
grep -c "CDR_DEV_ASSERT failed" $out/bounds_*.log
