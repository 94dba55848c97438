"""GPU tier, multi-GPU: row-sharded corpus over 2 ranks (NCCL all-gather of per-shard top-k +
K4 merge) returns exactly the single-corpus answer.  Skipped on a 1-GPU box; the 1-GPU tier
covers the same data path with logical shards (test_gpu_engine.py) and the CPU tier covers the
collective plumbing with gloo (test_dist_cpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, k, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from cadence_rag_b200.dist import ShardedSearcher, shard_range
    from cadence_rag_b200.store import DenseStore, SYNTH_QUERY_SEED, synth_rows_device
    first, cnt = shard_range(n_total, rank, world)
    store = DenseStore("chunks", cnt, dim=1024, device=rank)
    store.append_synthetic(cnt, first_row=first)
    store.finalize()
    searcher = ShardedSearcher(store, transport="nccl")
    qs = synth_rows_device(SYNTH_QUERY_SEED, 0, 130, 1024, device=rank)
    res = {}
    for mode in ("exact", "ann"):
        ids, sc, n = searcher.search(qs, k, mode=mode)
        torch.cuda.synchronize()
        res[mode] = (ids.cpu().numpy(), sc.cpu().numpy(), n.cpu().numpy())
    # K4p: the same exchange over NVLink peer memory (one push+merge kernel per rank) -- identical bits,
    # across many epochs (slot parity reuse), batches larger than the slot count (chunking), nq=1, small k
    peer = ShardedSearcher(store, transport="peer", max_nq=64, max_k=64)
    assert peer.transport == "peer"
    for rep in range(6):
        for nq_i, kk in ((130, k), (1, k), (64, 10), (65, k)):
            a = searcher.search(qs[:nq_i], kk, mode="exact")
            b = peer.search(qs[:nq_i], kk, mode="exact")
            torch.cuda.synchronize()
            for x, y in zip(a, b):
                assert torch.equal(x.view(torch.int64) if x.dtype == torch.float64 else x,
                                   y.view(torch.int64) if y.dtype == torch.float64 else y), (rep, nq_i, kk)
    pb = peer.search(qs, k, mode="ann")
    torch.cuda.synchronize()
    assert np.array_equal(pb[0].cpu().numpy(), res["ann"][0])
    # a filter that empties one rank's shard: short / empty local lists still merge
    allow, cnt_allowed = store.filter_bitmap(call_slots=[0, 1])          # rows 0..399 live on rank 0 only
    fa = searcher.search(qs[:5], k, allow=allow, mode="exact")
    fb = peer.search(qs[:5], k, allow=allow, mode="exact")
    torch.cuda.synchronize()
    assert torch.equal(fa[0], fb[0]) and torch.equal(fa[2], fb[2]) and int(fa[2][0]) == k
    assert (cnt_allowed == 400) == (rank == 0)
    dist.barrier()
    peer.close()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), e_ids=res["exact"][0], e_sc=res["exact"][1],
             a_ids=res["ann"][0], a_sc=res["ann"][1])
    if rank == 0:
        whole = DenseStore("chunks", n_total, dim=1024, device=0)
        whole.append_synthetic(n_total)
        whole.finalize()
        w_ids, w_sc, w_n = whole.search_exact(qs, k)
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, "whole.npz"), ids=w_ids.cpu().numpy(), sc=w_sc.cpu().numpy())
        whole.close()
    store.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_sharded_equals_whole(tmp_path):
    import torch.multiprocessing as mp
    n_total, k, world = 300_001, 50, 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, k, str(tmp_path)), nprocs=world, join=True)
    whole = np.load(tmp_path / "whole.npz")
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for r in (r0, r1):
        assert np.array_equal(r["e_ids"], whole["ids"])
        assert np.array_equal(r["e_sc"].view(np.uint64), whole["sc"].view(np.uint64))
        recall = np.mean([len(set(r["a_ids"][i]) & set(whole["ids"][i])) / k for i in range(whole["ids"].shape[0])])
        assert recall >= 0.999
    assert np.array_equal(r0["a_ids"], r1["a_ids"])       # every rank holds the identical merged result


def _crafted_lists(rank, world, nq=6, k=24):
    """Rank `rank`'s ordered (scores [nq,k], ids [nq,k], n [nq]) lists for the exchange test; ids are distinct across ranks."""
    rng = np.random.default_rng(1000 + rank)
    pool = np.array([1.0, 0.5, 0.5, 0.25, 0.0, -0.0, -0.25, np.inf, -np.inf, np.nan, np.nan, 0.125])
    sc = np.full((nq, k), 77.0); ids = np.full((nq, k), 5, dtype=np.int64); n = np.zeros(nq, dtype=np.int32)
    for q in range(nq):
        m = 0 if (q == 1 and rank == 0) else int(rng.integers(0, k + 1))
        vals = rng.choice(pool, size=m)
        own = (rng.permutation(4 * k)[:m] * world + rank).astype(np.int64)
        order = sorted(range(m), key=lambda i: (vals[i] != vals[i], -vals[i] if vals[i] == vals[i] else 0.0, own[i]))
        sc[q, :m] = vals[order]; ids[q, :m] = own[order]; n[q] = m
    return sc, ids, n


def _one_gpu_worker(rank, world, port, n_total, k, out_dir, pipeline, fused):
    """Two ranks on ONE device (gloo process group; the peer buffers are mapped with CUDA IPC within the device): the
    K4p exchange kernel of each rank really waits for the other process's push.  Kernels of the two processes are
    time-sliced by the driver, so a spinning exchange kernel is preempted for its peer -- slow, but the protocol and the
    merge are the ones the multi-GPU runs use."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CADENCE_SHARD_PIPELINE=str(pipeline),
                      CADENCE_PEER_FUSED=str(fused))
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cadence_rag_b200.dist import ShardedSearcher, shard_range
    from cadence_rag_b200.store import DenseStore, SYNTH_QUERY_SEED, synth_rows_device
    first, cnt = shard_range(n_total, rank, world)
    store = DenseStore("chunks", cnt, dim=1024, device=0)
    store.append_synthetic(cnt, first_row=first)
    store.finalize()
    peer = ShardedSearcher(store, transport="peer", max_nq=64, max_k=64)
    assert peer.transport == "peer" and peer.world == world
    qs = synth_rows_device(SYNTH_QUERY_SEED, 0, 70, 1024, device=0)
    out = {}
    # k60: candidate width 256 on the exact lane (not fusable: separate exchange launch); scan64: the widest fused form
    for name, nq_i, kk, mode in (("exact", 70, k, "exact"), ("one", 1, k, "exact"), ("k10", 64, 10, "exact"),
                                 ("shared", 40, k, "exact_shared"), ("ann", 70, k, "ann"), ("scan", 3, k, "scan_bf16"),
                                 ("k60", 5, 60, "exact"), ("scan64", 2, 64, "scan_bf16")):
        ids, sc, n = peer.search(qs[:nq_i], kk, mode="exact" if mode == "exact_shared" else mode, shared=mode == "exact_shared")
        torch.cuda.synchronize()
        out[name] = (ids.cpu().numpy(), sc.cpu().numpy(), n.cpu().numpy())
    allow, _cnt = store.filter_bitmap(call_slots=[0, 1])                  # rows 0..399: rank 0's shard only
    ids, sc, n = peer.search(qs[:5], k, allow=allow, mode="exact")
    torch.cuda.synchronize()
    out["filtered"] = (ids.cpu().numpy(), sc.cpu().numpy(), n.cpu().numpy())
    # the exchange on crafted ORDERED lists: ties across ranks (broken by id), NaN tails, signed zeros, infinities, short
    # and empty lists -- the merge works on integer order keys and must hand back every score's own bits
    c_sc, c_ids, c_n = _crafted_lists(rank, world)
    m_ids, m_sc, m_n = peer.peer.exchange_merge(torch.from_numpy(c_ids).cuda(), torch.from_numpy(c_sc).cuda(),
                                                torch.from_numpy(c_n).cuda(), c_sc.shape[1])
    torch.cuda.synchronize()
    out["crafted"] = (m_ids.cpu().numpy(), m_sc.cpu().numpy(), m_n.cpu().numpy())
    dist.barrier()
    peer.close()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **{f"{k_}_{part}": v[i] for k_, v in out.items()
                                                         for i, part in enumerate(("ids", "sc", "n"))})
    store.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("pipeline,fused,world", [(0, 1, 2), (0, 0, 2), (1, 1, 2), (0, 1, 3)])
def test_two_ranks_on_one_gpu_peer_exchange_equals_whole(tmp_path, pipeline, fused, world):
    """The multi-rank exchange on a ONE-GPU box: 2 processes share cuda:0, each owns half of the rows, and every
    merged result must carry the bits of the unsharded scan (exact lanes) or reach recall 0.999 against it (bf16 lanes)
    and be identical on both ranks.  pipeline=1: the chunked form of the step (finalize + exchange of chunk c on a side
    stream beside the scan of chunk c+1, CADENCE_SHARD_PIPELINE=1; off by default).  fused=1 (default): the scan lanes'
    finalize kernels end with the exchange (push, wait, merge in the CTA that ordered the list); fused=0: the separate K4p
    launch behind the lane (CADENCE_PEER_FUSED=0)."""
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    n_total, k = 60_001, 50
    mp.spawn(_one_gpu_worker, args=(world, _free_port(), n_total, k, str(tmp_path), pipeline, fused), nprocs=world, join=True)
    from cadence_rag_b200.store import DenseStore, SYNTH_QUERY_SEED, synth_rows_device
    whole = DenseStore("chunks", n_total, dim=1024, device=0)
    whole.append_synthetic(n_total)
    whole.finalize()
    qs = synth_rows_device(SYNTH_QUERY_SEED, 0, 70, 1024, device=0)
    r0 = np.load(tmp_path / "rank0.npz")
    for r in range(1, world):
        rr = np.load(tmp_path / f"rank{r}.npz")
        for key in r0.files:
            assert np.array_equal(r0[key], rr[key], equal_nan=rr[key].dtype.kind == "f"), key    # every rank holds the identical result
    for name, nq_i, kk in (("exact", 70, k), ("one", 1, k), ("k10", 64, 10), ("shared", 40, k), ("k60", 5, 60)):
        w_ids, w_sc, w_n = whole.search_exact(qs[:nq_i], kk)
        torch.cuda.synchronize()
        assert np.array_equal(r0[f"{name}_ids"], w_ids.cpu().numpy()), name
        assert np.array_equal(r0[f"{name}_sc"].view(np.uint64), w_sc.cpu().numpy().view(np.uint64)), name
        assert np.array_equal(r0[f"{name}_n"], w_n.cpu().numpy())
    for name, nq_i, kk in (("ann", 70, k), ("scan", 3, k), ("scan64", 2, 64)):
        w_ids = whole.search_exact(qs[:nq_i], kk)[0].cpu().numpy()
        got = r0[f"{name}_ids"]
        assert np.mean([len(set(got[i]) & set(w_ids[i])) / kk for i in range(nq_i)]) >= 0.999, name
    allow, cnt_allowed = whole.filter_bitmap(call_slots=[0, 1])
    f_ids, f_sc, f_n = whole.search_exact(qs[:5], k, allow)
    torch.cuda.synchronize()
    assert cnt_allowed == 400 and np.array_equal(r0["filtered_ids"], f_ids.cpu().numpy())
    assert np.array_equal(r0["filtered_sc"].view(np.uint64), f_sc.cpu().numpy().view(np.uint64))
    whole.close()
    from oracle import ports
    parts = [_crafted_lists(r, world) for r in range(world)]
    kk = parts[0][0].shape[1]
    want = ports.merge_topk(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), np.stack([p[2] for p in parts]), kk)
    for q, (w_ids, w_sc) in enumerate(want):
        m = len(w_ids)
        assert int(r0["crafted_n"][q]) == m, q
        assert r0["crafted_ids"][q, :m].tolist() == w_ids, q
        assert np.array_equal(r0["crafted_sc"][q, :m].view(np.uint64), np.array(w_sc, dtype=np.float64).view(np.uint64)), q
        assert (r0["crafted_ids"][q, m:] == -1).all()


# ----------------------------------------------------------------------------------------- sharded hybrid /retrieve
def _hybrid_corpus():
    """Deterministic 2-table corpus (same on every rank): rows, ids, calls, dates, tags, tech tokens, payload."""
    from datetime import datetime, timedelta, timezone
    from uuid import UUID
    sys.path.insert(0, ROOT)
    from oracle import cpu_oracle as orc
    rng = np.random.default_rng(23)
    t0 = datetime(2026, 1, 1, tzinfo=timezone.utc)
    tables = {}
    for name, rows, seed, key in (("chunks", 6000, 20260209, "chunk_id"), ("artifact_chunks", 600, 20260214, "artifact_chunk_id")):
        x = orc.synth_rows(seed, 0, rows)
        call_of_row = np.arange(rows) // (rows // 60)
        # calls are NOT in time order along the rows, so the tech lane's (started_at DESC, id ASC) interleaves shards
        hour_of_call = rng.permutation(60)
        zipf = np.minimum(rng.zipf(1.3, size=(rows, 3)) - 1, 99)
        ntok = rng.integers(0, 4, size=rows)
        valid = np.ones(rows, dtype=bool); valid[rng.choice(rows, 15, replace=False)] = False
        tables[name] = dict(x=x, ids=np.arange(1, rows + 1, dtype=np.int64) * 3, key=key, valid=valid,
                            call_ids=[UUID(int=int(c) + 1) for c in call_of_row],
                            started=[t0 + timedelta(hours=int(hour_of_call[c])) for c in call_of_row],
                            tags=[[f"t{c % 5}", f"u{c % 3}"] for c in call_of_row],
                            tokens=[[f"TOK-{z}" for z in zipf[r, : ntok[r]]] for r in range(rows)],
                            payload=[{"text": f"{name} row {r}"} for r in range(rows)])
    return tables, t0


def _build_engine(tables, lo_frac, hi_frac, device):
    from cadence_rag_b200.lexical import TechTokenIndex
    from cadence_rag_b200.retrieve import DenseEngine
    from cadence_rag_b200.store import DenseStore
    from uuid import UUID
    eng = DenseEngine()
    for name, t in tables.items():
        n = len(t["ids"])
        lo, hi = int(n * lo_frac), int(n * hi_frac)
        store = DenseStore(name, max(hi - lo, 1), dim=1024, device=device)
        store.append(t["x"][lo:hi], ids=t["ids"][lo:hi], call_ids=t["call_ids"][lo:hi], call_started_at=t["started"][lo:hi],
                     call_tags=t["tags"][lo:hi], valid=t["valid"][lo:hi], payload=t["payload"][lo:hi])
        store.finalize()
        index = TechTokenIndex()
        for r in range(lo, hi):
            index.add_row(r - lo, t["tokens"][r])
        eng.register(store, index)
    for c in range(60):
        eng.register_call(UUID(int=c + 1), external_id=f"ext-{c // 2}", external_source="crm" if c % 2 else None)
    return eng


def _hybrid_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import json
    from datetime import timedelta
    from uuid import UUID
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from cadence_rag_b200 import embeddings, retrieve
    from cadence_rag_b200.config import settings
    from cadence_rag_b200.retrieve import RetrieveFilters
    from cadence_rag_b200.sharded import ShardedEngine, sharded_retrieve_ids
    settings.embeddings_dim = 1024
    settings.cadence_gpu_device = rank
    embeddings.set_embedder(embeddings.SyntheticEmbedder(seed=20260210, dim=1024))
    tables, t0 = _hybrid_corpus()
    shard = ShardedEngine(_build_engine(tables, rank / world, (rank + 1) / world, rank))
    assert all(s.transport == "peer" for s in shard.searchers.values())
    bm25 = [{"chunk_id": 3 * i} for i in (5, 4100, 77, 5999)]
    bm25a = [{"artifact_chunk_id": 3 * i} for i in (3, 450)]
    requests = [("why did TOK-1 fail with TOK-3 on 10.0.0.1", None, bm25, bm25a),
                ("TOK-0 status", RetrieveFilters(call_ids=[UUID(int=c + 1) for c in range(25, 35)]), [], []),     # spans both shards
                ("TOK-2 and TOK-5", RetrieveFilters(call_tags=["t2"]), bm25, []),
                ("no tokens here", RetrieveFilters(date_from=t0 + timedelta(hours=30)), [], bm25a),
                ("TOK-4 TOK-9", RetrieveFilters(external_id="ext-7"), bm25, bm25a),                                  # lives on one shard
                ("TOK-4", RetrieveFilters(external_id="missing"), bm25, []),
                ("   ", None, [], [])]
    got = [sharded_retrieve_ids(shard, q, f, bm25_chunks=b, bm25_artifacts=a, debug=True) for q, f, b, a in requests]
    # dense lane disabled on every rank -> lexical-only fusion
    embeddings.set_embedder(None)
    settings.embeddings_base_url = ""
    got.append(sharded_retrieve_ids(shard, "TOK-1 TOK-3", None, bm25_chunks=bm25, debug=True))
    with open(os.path.join(out_dir, f"sharded{rank}.json"), "w") as f:
        json.dump(got, f, default=str)
    if rank == 0:
        embeddings.set_embedder(embeddings.SyntheticEmbedder(seed=20260210, dim=1024))
        whole = _build_engine(tables, 0.0, 1.0, 0)
        want = [retrieve.retrieve_ids(whole, q, f, bm25_chunks=b, bm25_artifacts=a, debug=True) for q, f, b, a in requests]
        embeddings.set_embedder(None)
        want.append(retrieve.retrieve_ids(whole, "TOK-1 TOK-3", None, bm25_chunks=bm25, debug=True))
        with open(os.path.join(out_dir, "whole.json"), "w") as f:
            json.dump(want, f, default=str)
        for st in whole.stores.values():
            st.close()
    dist.barrier()
    shard.close()
    for st in shard.local.stores.values():
        st.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_sharded_hybrid_retrieve_equals_whole(tmp_path):
    """cadence_rag_b200.sharded.sharded_retrieve_ids over 2 row shards == retrieve_ids over the whole corpus: lanes,
    COUNT(*), planner modes, fused ranks (bit-exact RRF) and the ids_only order, identical on both ranks."""
    import json
    import torch.multiprocessing as mp
    mp.spawn(_hybrid_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    want = json.load(open(tmp_path / "whole.json"))
    for rank in (0, 1):
        got = json.load(open(tmp_path / f"sharded{rank}.json"))
        assert len(got) == len(want) == 8
        for i, (g, w) in enumerate(zip(got, want)):
            assert g == w, (rank, i)
    assert len(want[0]["retrieved_ids"]) > 50 and want[-1]["debug"]["dense"]["enabled"] is False
