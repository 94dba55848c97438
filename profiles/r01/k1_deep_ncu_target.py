"""ncu target: one 64-query shared-read call (deep kernel, 8 query groups on 18 CTAs each) over 1M x 1024 fp32 rows."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from cadence_rag_b200.store import DenseStore, SYNTH_QUERY_SEED, synth_rows_device
s = DenseStore("chunks", 1_000_000, dim=1024, device=0, fp32=True, bf16=False)
s.append_synthetic(1_000_000); s.finalize()
q = synth_rows_device(SYNTH_QUERY_SEED, 0, 64, 1024, device=0)
for _ in range(3):
    s.search_exact(q, 50, shared=True)
torch.cuda.synchronize()
print("ok")
