#!/bin/bash
# r02 experiment 32 (1 GPU): the A/B paths stay tested -- GPU tier with CADENCE_PDL=0, with CADENCE_PEER_FUSED=0 +
# CADENCE_TECH_CLUSTER=0 + CADENCE_FIN_CLUSTER=0, and on the bounds build of the final code
cd $GRAFT_REPO_ROOT
CADENCE_PDL=0 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
CADENCE_PEER_FUSED=0 CADENCE_TECH_CLUSTER=0 CADENCE_FIN_CLUSTER=0 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
CADENCE_DENSE_LIB=$PWD/build/ab/bounds/libcadence_dense.so python -m pytest tests -m gpu -x -q 2>&1 | tail -2
LD_LIBRARY_PATH=$PWD/build/ab/bounds ./build/sanitize_driver 2>&1 | tail -2
