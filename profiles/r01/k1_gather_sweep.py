"""K1 gather launch vs full scan as a function of filter selectivity (run twice: CADENCE_K1_GATHER_DIV=1 forces the
gather launch for every filter, CADENCE_K1_GATHER=0 forces the full scan).  1M x 1024 fp32, one query, k=50."""
import os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cadence_rag_b200.store import DenseStore, SYNTH_QUERY_SEED, synth_rows_device

n = 1_000_000
s = DenseStore("chunks", n, dim=1024, device=0, fp32=True, bf16=False)
s.append_synthetic(n); s.finalize()
q = synth_rows_device(SYNTH_QUERY_SEED, 0, 64, 1024, device=0)
rng = np.random.default_rng(0)
out = {}
for name, frac, block in [("1/64 random", 1 / 64, 1), ("1/16 random", 1 / 16, 1), ("1/8 random", 1 / 8, 1), ("1/4 random", 1 / 4, 1),
                          ("1/2 random", 1 / 2, 1), ("1/16 blocks of 200", 1 / 16, 200), ("1/4 blocks of 200", 1 / 4, 200)]:
    keep = np.zeros(n, dtype=bool)
    if block == 1:
        keep[rng.choice(n, int(n * frac), replace=False)] = True
    else:
        nb = n // block
        for b in rng.choice(nb, int(nb * frac), replace=False):
            keep[b * block:(b + 1) * block] = True
    bits = np.packbits(keep.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1)
    allow = torch.from_numpy(bits.view(np.int32).copy()).cuda()
    for i in range(5):
        s.search_exact(q[i:i + 1], 50, allow)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(40):
        s.search_exact(q[i:i + 1], 50, allow)
    b.record(); torch.cuda.synchronize()
    out[name] = round(a.elapsed_time(b) / 40, 4)
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("CADENCE_K1")}, "ms_per_query": out}))
