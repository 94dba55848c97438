import sys, os
sys.path.insert(0, os.getcwd())
import torch
from cadence_rag_b200.store import DenseStore, SYNTH_QUERY_SEED, synth_rows_device
n = 1_000_000
s = DenseStore("chunks", n, dim=1024, device=0, fp32=True, bf16=False)
s.append_synthetic(n); s.finalize()
q = synth_rows_device(SYNTH_QUERY_SEED, 0, 48, 1024, device=0)
for i in range(3):
    s.search_exact(q, 50, shared=True)
torch.cuda.synchronize()
