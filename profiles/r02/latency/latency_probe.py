"""Single-query latency of the (sharded) scan lanes, device-timed (CUDA events around the one call a client makes).
Run alone (N = 1) or under torch.distributed.run (N ranks, 1 M rows sharded).  Env A/B: CADENCE_PEER_FUSED, CADENCE_PDL.
Prints one JSON line on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.environ.get("PROBE_PKG_ROOT", ROOT))      # (A/B of the Python side: another copy of the package)
import torch
import torch.distributed as dist

from cadence_rag_b200.dist import ShardedSearcher, shard_range
from cadence_rag_b200.store import DenseStore, SYNTH_QUERY_SEED, synth_rows_device

ROWS, DIM, K, ITERS = int(os.environ.get("PROBE_ROWS", 1_000_000)), 1024, 50, int(os.environ.get("PROBE_ITERS", 200))
world = int(os.environ.get("WORLD_SIZE", 1))
rank = int(os.environ.get("RANK", 0))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
first, count = shard_range(ROWS, rank, world)
store = DenseStore("chunks", count, dim=DIM, device=local, fp32=True, bf16=True)
store.append_synthetic(count, first_row=first)
store.finalize()
searcher = ShardedSearcher(store)
qs = synth_rows_device(SYNTH_QUERY_SEED, 0, 64, DIM, device=local)
out = {"n_gpus": world, "rows": ROWS, "k": K, "iters": ITERS, "exchange": searcher.transport,
       "env": {k: os.environ.get(k) for k in ("CADENCE_PEER_FUSED", "CADENCE_PDL")}}
for mode in ("exact", "scan_bf16"):
    single = [qs[i:i + 1].contiguous() for i in range(64)]
    for i in range(20):
        searcher.search(single[i], K, mode=mode)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    lat = []
    for i in range(ITERS):
        q = single[i % 64]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); searcher.search(q, K, mode=mode); b.record()
        torch.cuda.synchronize()
        lat.append(a.elapsed_time(b))
    lat.sort()
    t = torch.tensor([lat[len(lat) // 2], lat[0], lat[len(lat) // 10]], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[mode] = {"p50_ms": float(t[0]), "min_ms": float(t[1]), "p10_ms": float(t[2])}
    # a batch of 64 (throughput form), for the step tail
    for i in range(3):
        searcher.search(qs, K, mode=mode)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(10):
        searcher.search(qs, K, mode=mode)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 10], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[mode]["batch64_ms"] = float(t[0])
if rank == 0:
    print(json.dumps(out))
searcher.close()
store.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
