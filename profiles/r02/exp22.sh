#!/bin/bash
# r02 experiment 22 (1 GPU): is the latency finalize bound by cold instruction fetch?  The probe build runs it twice.
cd $GRAFT_REPO_ROOT
CADENCE_FIN_TWICE=0 PROBE_ITERS=6 CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_fintiming.so python profiles/r02/latency/latency_probe.py > gpurun_out/r02_exp22_fintwice_1gpu.log 2>&1
grep FIN gpurun_out/r02_exp22_fintwice_1gpu.log | tail -8
