"""Small batches on the bf16-row scan lane: ms per call for nq queries over 1 M x 1024 rows, device-timed, beside the shared
fp32 scan and the tensor-core lane for the same batch.  CADENCE_BF16_SHARE = 0 / 2 / 4 / unset selects the sharing form."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cadence_rag_b200.store import DenseStore, SYNTH_QUERY_SEED, synth_rows_device

s = DenseStore("chunks", 1_000_000, dim=1024, device=0, fp32=True, bf16=True)
s.append_synthetic(1_000_000); s.finalize()
q = synth_rows_device(SYNTH_QUERY_SEED, 0, 64, 1024, device=0)
out = {"share": os.environ.get("CADENCE_BF16_SHARE"), "rows": 1_000_000}

def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

for nq in (1, 2, 3, 4, 6, 8, 16, 64):
    qq = q[:nq].contiguous()
    rec = {"scan_bf16_ms": timed(lambda: s.search_scan_bf16(qq, 50))}
    if os.environ.get("CADENCE_BF16_SHARE") is None:
        rec["exact_shared_fp32_ms"] = timed(lambda: s.search_exact(qq, 50, shared=True))
        rec["batch_bf16_ms"] = timed(lambda: s.search_batch(qq, 50))
    out[str(nq)] = rec
print(json.dumps(out))
