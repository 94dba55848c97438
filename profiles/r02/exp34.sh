#!/bin/bash
# r02 experiment 34 (1 GPU): K6 with four words per trip -- filter tests, then ms per call against the library of commit dc19172
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_engine.py -x -q -k "filter or hybrid or tech or growth or concurrent or tag" 2>&1 | tail -3
for i in 1 2; do
CADENCE_DENSE_LIB=$GRAFT_REPO_ROOT/build/ab/libcadence_dense_head_dc19172.so python profiles/r02/filter_probe.py | sed 's/^/dc19172 /'
python profiles/r02/filter_probe.py | sed 's/^/new     /'
done
