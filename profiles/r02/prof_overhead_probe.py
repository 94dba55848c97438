"""What do the per-launch CUDA events of cdr_prof cost a K2 step?  10M x 1024 bf16, 1024 queries per step."""
import json, os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import torch
from cadence_rag_b200 import _ffi
from cadence_rag_b200.store import DenseStore, SYNTH_QUERY_SEED, synth_rows_device
rows, nq, dim = 10_000_000, 1024, 1024
store = DenseStore("chunks", rows, dim=dim, device=0, fp32=True, bf16=True)
store.append_synthetic(rows); store.finalize()
q = synth_rows_device(SYNTH_QUERY_SEED, 0, 8 * nq, dim, device=0).view(8, nq, dim)
def run(steps, prof):
    _ffi.lib().cdr_prof_enable(1 if prof else 0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for s in range(steps):
        store.search_batch(q[s % 8], 50)
    b.record(); torch.cuda.synchronize()
    _ffi.lib().cdr_prof_enable(0)
    return a.elapsed_time(b) / steps
run(3, False)
out = []
for rep in range(3):
    out.append({"prof_off_ms_per_step": run(12, False), "prof_on_ms_per_step": run(12, True)})
print(json.dumps(out))
