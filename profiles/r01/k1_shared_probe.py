"""Shared reads on the exact lane: ms per call and queries/s for batches of 1..64 queries over 1M x 1024 fp32 rows.
CADENCE_K1_DEEP=0 keeps the 3-queries-in-registers kernel for batches > 3 (A/B)."""
import json, os, sys
sys.path.insert(0, os.getcwd())
import torch
from cadence_rag_b200.store import DenseStore, SYNTH_QUERY_SEED, synth_rows_device
n = int(os.environ.get("PROBE_ROWS", 1_000_000))
s = DenseStore("chunks", n, dim=1024, device=0, fp32=True, bf16=False)
s.append_synthetic(n); s.finalize()
q = synth_rows_device(SYNTH_QUERY_SEED, 0, 64, 1024, device=0)
out = {}
for nq in [int(v) for v in os.environ.get("PROBE_NQ", "1,3,4,8,16,24,48,64").split(",")]:
    qq = q[:nq].contiguous()
    for shared in (False, True):
        for _ in range(3):
            s.search_exact(qq, 50, shared=shared)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            s.search_exact(qq, 50, shared=shared)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        out[f"nq={nq} shared={int(shared)}"] = {"ms": round(ms, 4), "qps": round(nq / ms * 1e3, 1)}
print(json.dumps({"deep": os.environ.get("CADENCE_K1_DEEP", "1"), "rows": n, "results": out}))
