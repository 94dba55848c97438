"""CPU tier: the independent parity oracle that bench.py runs in every record (plain PyTorch fp32 matmul + fp64
re-score) is itself checked against the C oracle -- on one process and on two gloo ranks (the multi-rank merge the
N > 1 bench lines use).  No GPU, no engine code."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import cpu_oracle as orc  # noqa: E402

SEED_C, SEED_Q = 20260209, 20260210


class _Ctx:
    """bench.Ctx without CUDA: what oracle_topk_torch needs."""

    def __init__(self, world=1, rank=0, dist=None):
        self.torch, self.dist, self.world, self.rank = torch, dist, world, rank


def _chunks(first, count, chunk):
    for off in range(0, count, chunk):
        m = min(chunk, count - off)
        yield first + off, torch.from_numpy(orc.synth_rows(SEED_C, first + off, m))


def test_torch_oracle_equals_c_oracle_single_process():
    n, k = 30_000, 50
    x = orc.synth_rows(SEED_C, 0, n)
    q = orc.synth_rows(SEED_Q, 0, 6)
    ids, sc = bench.oracle_topk_torch(_Ctx(), _chunks(0, n, 7_000), torch.from_numpy(q), k)
    for i in range(6):
        w_ids, w_sc = orc.exact_scan(q[i], x, k, variant=orc.VARIANT_F64)
        assert ids[i].tolist() == w_ids.tolist()
        assert np.allclose(sc[i].numpy(), w_sc, rtol=1e-12, atol=0)
    # compare_lists is what the bench asserts on
    recall, ident, rel = bench.compare_lists(torch, ids, sc, ids.clone(), sc.clone(), k)
    assert (recall, ident, rel) == (1.0, 1.0, 0.0)
    worse = ids.clone(); worse[0, 0] = -5
    recall, ident, _ = bench.compare_lists(torch, worse, sc, ids, sc, k)
    assert recall < 1.0 and ident < 1.0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, n, k, out):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        first, count = bench_shard(n, rank, world)
        q = torch.from_numpy(orc.synth_rows(SEED_Q, 100, 4))
        ids, sc = bench.oracle_topk_torch(_Ctx(world, rank, dist), _chunks(first, count, 5_000), q, k)
        out[rank] = (ids.numpy().copy(), sc.numpy().copy())
    finally:
        dist.destroy_process_group()


def bench_shard(n, rank, world):
    per = (n + world - 1) // world
    first = min(n, rank * per)
    return first, max(0, min(n, first + per) - first)


def test_torch_oracle_two_gloo_ranks_merge_to_the_unsharded_answer():
    import torch.multiprocessing as mp
    n, k, world = 20_001, 50, 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, n, k, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    x = orc.synth_rows(SEED_C, 0, n)
    q = orc.synth_rows(SEED_Q, 100, 4)
    for r in range(world):
        ids, sc = out[r]
        for i in range(4):
            w_ids, w_sc = orc.exact_scan(q[i], x, k, variant=orc.VARIANT_F64)
            assert ids[i].tolist() == w_ids.tolist(), (r, i)
            assert np.allclose(sc[i], w_sc, rtol=1e-12, atol=0)
