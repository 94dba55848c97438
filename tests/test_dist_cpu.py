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


# ----------------------------------------------------------------------------------------- sharded hybrid facade (host logic)
def _sharded_worker(rank, world, port, out_dir):
    """cadence_rag_b200.sharded over 2 gloo ranks with the GPU pieces replaced by the oracle: what is exercised is the
    cross-rank plumbing -- embedding broadcast from rank 0, all-reduced COUNT(*), tech-lane merge by
    (call_started_at DESC, id ASC), row dicts from the owning rank, identical responses on every rank."""
    sys.path.insert(0, ROOT)
    import contextlib
    import json
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cadence_rag_b200 import embeddings, retrieve as R, sharded
    from cadence_rag_b200.config import settings
    from oracle import cpu_oracle as orc
    from oracle import ports
    settings.embeddings_dim = 64
    n, k_dim = 400, 64
    x = orc.synth_rows(20260209, 0, n, k_dim)
    ids_all = np.arange(1, n + 1, dtype=np.int64) * 2
    rng = np.random.default_rng(4)
    started_all = rng.permutation(n).astype(np.int64) // 7            # ties in call_started_at, interleaved across shards
    tokens_all = [["TOK-1"] if r % 3 == 0 else (["TOK-2"] if r % 5 == 0 else []) for r in range(n)]
    lo, hi = (0, 230) if rank == 0 else (230, n)                      # uneven shards

    class _Store:
        table_name, key_field, dim, device = "chunks", "chunk_id", k_dim, 0
        call_ids_by_slot = [f"call-{c}" for c in range(40)]
        payload = {int(i): {"text": f"row {i}"} for i in ids_all[lo:hi]}

        def host_columns(self):
            return {"ids": ids_all[lo:hi], "call_slot": (np.arange(lo, hi) % 40).astype(np.int32), "started_at": started_all[lo:hi]}

    class _Local:
        stores = {"chunks": _Store()}
        external_ids = {}

        @contextlib.contextmanager
        def connect(self):
            conn = type("C", (), {})()
            conn.engine = self
            conn.store = lambda t: self.stores[t]
            yield conn

    class _Searcher:                                                 # the dense lane + exchange, restated with the oracle
        transport = "fake"

        def search(self, qd, limit, allow, mode="exact"):
            g_ids, g_sc = orc.exact_scan(qd[0].numpy(), x, limit, ids=ids_all)
            ids = torch.full((1, limit), -1, dtype=torch.int64); sc = torch.full((1, limit), float("nan"), dtype=torch.float64)
            ids[0, :len(g_ids)] = torch.from_numpy(g_ids); sc[0, :len(g_sc)] = torch.from_numpy(g_sc)
            return ids, sc, torch.tensor([len(g_ids)], dtype=torch.int32)

        def close(self):
            pass

    def fake_fetch_tech(conn, table, tokens, filters, call_ids, limit):
        rows = [r for r in range(lo, hi) if set(tokens) & set(tokens_all[r])]
        rows.sort(key=lambda r: (-int(started_all[r]), int(ids_all[r])))
        return [{"chunk_id": int(ids_all[r]), "call_id": f"call-{r % 40}", "text": f"row {ids_all[r]}"} for r in rows[:limit]]

    R._fetch_tech = fake_fetch_tech
    R._filter_bitmap = lambda conn, table, filters, call_ids: (None, hi - lo)
    R._rrf_merge = lambda lanes, key, k=60: ports.rrf_merge(lanes, key, k)
    eng = sharded.ShardedEngine.__new__(sharded.ShardedEngine)
    eng.local, eng.group, eng.world, eng.rank, eng.searchers = _Local(), None, world, rank, {"chunks": _Searcher()}
    calls = []

    def embedder(texts):
        calls.append(texts)
        return embeddings.EmbeddingResult(vectors=[orc.synth_rows(20260210, 5, 1, k_dim)[0].astype(float).tolist()], model="m")
    embeddings.set_embedder(embedder)
    out = sharded.sharded_retrieve_ids(eng, "what about TOK-1 and TOK-2", None, bm25_chunks=[{"chunk_id": 8}, {"chunk_id": 700}], debug=True)
    assert (len(calls) == 1) == (rank == 0)                          # rank 0 alone talks to the embedding service
    with open(os.path.join(out_dir, f"resp{rank}.json"), "w") as f:
        json.dump(out, f, default=str)
    if rank == 0:                                                    # expectation from the unsharded restatement
        q = orc.synth_rows(20260210, 5, 1, k_dim)[0]
        d_ids, d_sc = orc.exact_scan(q, x, 50, ids=ids_all)
        rows = [r for r in range(n) if {"TOK-1", "TOK-2"} & set(tokens_all[r])]
        rows.sort(key=lambda r: (-int(started_all[r]), int(ids_all[r])))
        t_ids = [int(ids_all[r]) for r in rows[:50]]
        fused = ports.rrf_merge({"bm25": [{"chunk_id": 8}, {"chunk_id": 700}], "tech_tokens": [{"chunk_id": i} for i in t_ids],
                                 "dense": [{"chunk_id": int(i)} for i in d_ids]}, "chunk_id")
        with open(os.path.join(out_dir, "want.json"), "w") as f:
            json.dump({"tech": t_ids, "dense": [int(i) for i in d_ids], "fused": [[r["chunk_id"], sorted(h), s_] for r, h, s_ in fused],
                       "count": n}, f)
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_sharded_hybrid_plumbing(tmp_path):
    import json
    mp.spawn(_sharded_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    want = json.load(open(tmp_path / "want.json"))
    r0, r1 = json.load(open(tmp_path / "resp0.json")), json.load(open(tmp_path / "resp1.json"))
    assert r0 == r1
    dbg = r0["debug"]
    assert [e["chunk_id"] for e in dbg["lanes"]["chunks"]["tech_tokens"]] == want["tech"]
    assert [e["chunk_id"] for e in dbg["lanes"]["chunks"]["dense"]] == want["dense"]
    assert dbg["fused"]["chunks"] == want["fused"]
    assert dbg["dense"]["candidate_rows"]["chunks"] == want["count"] and dbg["dense"]["modes"]["chunks"] == "ann"
    assert r0["retrieved_ids"][0].startswith("chunk:")
