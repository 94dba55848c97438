"""ncu target: the exact scan at BASELINE configs[1] -- one 64-query launch (CTAs persistent over the queries) and one
single-query launch over the 1M x 1024 fp32 corpus, plus one bf16-row scan (FHFMA) -- after a warm-up of each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cadence_rag_b200.store import DenseStore, SYNTH_QUERY_SEED, synth_rows_device

s = DenseStore("chunks", 1_000_000, dim=1024, device=0, fp32=True, bf16=True)
s.append_synthetic(1_000_000); s.finalize()
q = synth_rows_device(SYNTH_QUERY_SEED, 0, 256, 1024, device=0)
for i in range(2):                       # warm-up: launches 0-5 of exact_scan_kernel
    s.search_exact(q[64 * i:64 * i + 64], 50); s.search_exact(q[i:i + 1], 50); s.search_scan_bf16(q[i:i + 1], 50)
torch.cuda.synchronize()
s.search_exact(q[128:192], 50)           # launch 6: 64 queries
s.search_exact(q[200:201], 50)           # launch 7: one query
s.search_scan_bf16(q[201:202], 50)       # launch 8: one query over the bf16 rows
torch.cuda.synchronize()
print("done")
