"""K2 over a small corpus (1M x 1024 bf16+fp32, 64 queries): per-call time and (under ncu) the launch list."""
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
from cadence_rag_b200.store import DenseStore, SYNTH_QUERY_SEED, synth_rows_device
n = int(os.environ.get("PROBE_ROWS", 1_000_000))
s = DenseStore("chunks", n, dim=1024, device=0)
s.append_synthetic(n); s.finalize()
out = {}
for nq in [int(v) for v in os.environ.get("PROBE_NQ", "16,64,256").split(",")]:
    q = synth_rows_device(SYNTH_QUERY_SEED, 0, nq, 1024, device=0)
    for _ in range(3):
        s.search_batch(q, 50)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        s.search_batch(q, 50)
    b.record(); torch.cuda.synchronize()
    out[f"nq={nq}"] = round(a.elapsed_time(b) / 10, 4)
print(json.dumps({"rows": n, "ms_per_call": out}))
