"""CPU tier: the N>1 host path with world_size-2 gloo processes (no GPU): shard ranges, the
packed single all-gather, and that the gathered lists merge (checked with the oracle port) to the
unsharded oracle answer."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, k, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cadence_rag_b200 import dist as cdist
    from oracle import cpu_oracle as orc
    from oracle import ports
    first, cnt = cdist.shard_range(n_total, rank, world)
    x = orc.synth_rows(20260209, first, cnt)
    qs = orc.synth_rows(20260210, 0, 3)
    ids = np.full((3, k), -1, dtype=np.int64); sc = np.full((3, k), np.nan); n = np.zeros(3, dtype=np.int32)
    for i in range(3):   # the oracle stands in for the local GPU search of this rank's shard
        li, ls = orc.exact_scan(qs[i], x, k, ids=np.arange(first + 1, first + cnt + 1))
        ids[i, :len(li)] = li; sc[i, :len(li)] = ls; n[i] = len(li)
    g_sc, g_id, g_n = cdist.gather_shard_results(torch.from_numpy(ids), torch.from_numpy(sc), torch.from_numpy(n))
    assert g_sc.shape == (world, 3, k) and g_id.dtype == torch.int64 and g_n.dtype == torch.int32
    assert np.array_equal(g_id[rank].numpy(), ids) and np.array_equal(g_n[rank].numpy(), n)
    assert np.array_equal(g_sc[rank].numpy().view(np.uint64), sc.view(np.uint64))   # bit views survive packing
    # transport selection: gloo ranks cannot map peer memory -> the all-gather transport, loudly for "peer"
    class _Store:
        device = 0
    searcher = cdist.ShardedSearcher(_Store(), transport="auto")
    assert searcher.transport == "nccl" and searcher.peer is None and searcher.world == world
    try:
        cdist.ShardedSearcher(_Store(), transport="peer")
        raise AssertionError("peer transport must refuse non-CUDA ranks")
    except cdist._ffi.DenseEngineError:
        pass
    merged = ports.merge_topk(g_sc.numpy(), g_id.numpy(), g_n.numpy(), k)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array([m[0] for m in merged]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_rows():
    from cadence_rag_b200.dist import shard_range
    for n, world in [(10, 3), (100_000_000, 8), (7, 8), (0, 2), (16, 4)]:
        spans = [shard_range(n, r, world) for r in range(world)]
        assert sum(c for _, c in spans) == n
        pos = 0
        for first, c in spans:
            assert first == pos or c == 0
            pos += c
    assert shard_range(100_000_000, 7, 8) == (87_500_000, 12_500_000)


def test_merge_requires_cuda():
    from cadence_rag_b200 import DenseEngineError
    from cadence_rag_b200.dist import merge_shard_results
    z = torch.zeros((2, 1, 4), dtype=torch.float64)
    with pytest.raises(DenseEngineError):
        merge_shard_results(z, z.long(), torch.zeros((2, 1), dtype=torch.int32), 4)


def test_world2_gloo_gather_and_merge(tmp_path):
    n_total, k, world = 5001, 20, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, k, str(tmp_path)), nprocs=world, join=True)
    from oracle import cpu_oracle as orc
    x = orc.synth_rows(20260209, 0, n_total)
    qs = orc.synth_rows(20260210, 0, 3)
    want = np.array([orc.exact_scan(qs[i], x, k)[0] for i in range(3)])
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"rank{r}.npy"), want)
