"""K6 timing: the allow bitmap + COUNT(*) of a 10-call filter (and a date + tag filter) over 1 M rows, ms per call through
DenseStore.filter_bitmap (host call: H2D of the call bitmap, kernel, D2H of the count, one sync)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cadence_rag_b200.store import DenseStore
s = DenseStore("chunks", 1_000_000, dim=1024, device=0, fp32=True, bf16=False)
s.append_synthetic(1_000_000); s.finalize()
out = {}
for name, kw in (("10_calls", dict(call_slots=list(range(10)))), ("calls_every_3rd", dict(call_slots=list(range(0, 5000, 3))))):
    for _ in range(5):
        s.filter_bitmap(**kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200):
        allow, cnt = s.filter_bitmap(**kw)
    torch.cuda.synchronize()
    out[name] = {"ms_per_call": (time.perf_counter() - t0) / 200 * 1e3, "count": int(cnt)}
print(json.dumps(out))
