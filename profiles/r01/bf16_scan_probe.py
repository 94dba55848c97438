"""Single-query "ann" lane (cdr_search_scan_bf16) vs the exact fp32 scan: kernel-level ms per query and recall."""
import json, os, sys
sys.path.insert(0, os.getcwd())
import torch
from cadence_rag_b200.store import DenseStore, SYNTH_QUERY_SEED, synth_rows_device
n = int(os.environ.get("PROBE_ROWS", 1_000_000))
s = DenseStore("chunks", n, dim=1024, device=0)
s.append_synthetic(n); s.finalize()
q = synth_rows_device(SYNTH_QUERY_SEED, 0, 64, 1024, device=0)
out = {"rows": n}
for name, fn in (("exact_f32", s.search_exact), ("scan_bf16", s.search_scan_bf16)):
    for i in range(5):
        fn(q[i:i + 1], 50)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(64):
        fn(q[i:i + 1], 50)
    b.record(); torch.cuda.synchronize()
    out[name + "_ms_per_query"] = round(a.elapsed_time(b) / 64, 4)
e = s.search_exact(q, 50); g = s.search_scan_bf16(q, 50); torch.cuda.synchronize()
out["recall_at_50"] = sum(len(set(e[0][i].tolist()) & set(g[0][i].tolist())) for i in range(64)) / (64 * 50)
out["identical_lists"] = sum(int(torch.equal(e[0][i], g[0][i])) for i in range(64))
print(json.dumps(out))
