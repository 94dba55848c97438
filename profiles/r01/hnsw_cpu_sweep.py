"""CPU-only: the restated HNSW baseline (oracle/hnsw_baseline.cc; m=16, ef_construction=64) over larger samples of the
synthetic corpus than bench.py can afford in-run: build time, queries/s (all cores / one thread) and recall@50 vs the
exact oracle at ef_search = 80 (the reference's setting) and larger beams.  Run in the build container."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cpu_oracle as orc

out = []
for n in [int(a) for a in sys.argv[1:]] or [10_000, 50_000]:
    x = orc.synth_rows(20260209, 0, n)
    qs = orc.synth_rows(20260210, 20_000_000, 256)
    t0 = time.perf_counter(); index = orc.HnswBaseline(x); build = time.perf_counter() - t0
    truth = [set((orc.exact_scan(qs[i], x, 50)[0] - 1).tolist()) for i in range(64)]
    row = {"rows": n, "build_seconds": round(build, 1), "cores": os.cpu_count()}
    for ef in (80, 200, 800):
        t0 = time.perf_counter(); rows, _, _ = index.search(qs, 50, ef); dt = time.perf_counter() - t0
        t1 = time.perf_counter(); index.search(qs[:32], 50, ef, nthreads=1); d1 = time.perf_counter() - t1
        rec = float(np.mean([len(truth[i] & set(rows[i].tolist())) / 50 for i in range(64)]))
        row[f"ef_search_{ef}"] = {"queries_per_s_all_cores": round(256 / dt, 1), "queries_per_s_one_thread": round(32 / d1, 1), "recall_at_50": round(rec, 4)}
    out.append(row); print(json.dumps(row), flush=True)
    index.close()
