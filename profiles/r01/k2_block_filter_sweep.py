"""K2 (batched bf16 tensor-core lane) under block-structured filters: tiles without an allowed row are skipped.
10M x 1024 bf16 (+fp32), 1024 queries per batch, k=50; date_from keeps the newest `frac` of the (time-ordered) rows."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cadence_rag_b200.store import DenseStore, SYNTH_CALL_PERIOD_US, SYNTH_QUERY_SEED, SYNTH_T0_US, synth_rows_device

n = 10_000_000
s = DenseStore("chunks", n, dim=1024, device=0, fp32=True, bf16=True)
s.append_synthetic(n); s.finalize()
q = synth_rows_device(SYNTH_QUERY_SEED, 0, 1024, 1024, device=0)
calls = n // 200
out = {}
for frac in (1.0, 0.5, 0.3, 0.1, 0.03):
    allow = None
    if frac < 1.0:
        allow, count = s.filter_bitmap(date_from=SYNTH_T0_US + int(calls * (1 - frac)) * SYNTH_CALL_PERIOD_US)
    for _ in range(2):
        s.search_batch(q, 50, allow)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        r = s.search_batch(q, 50, allow)
    b.record(); torch.cuda.synchronize()
    e = s.search_exact(q[:16].contiguous(), 50, allow)
    torch.cuda.synchronize()
    rec = sum(len(set(r[0][i].tolist()) & set(e[0][i].tolist())) for i in range(16)) / (16 * 50)
    out[f"newest {frac:.2f} of the rows"] = {"ms_per_batch": round(a.elapsed_time(b) / 5, 3), "recall_vs_exact_lane": rec}
print(json.dumps(out))
