"""Hybrid /retrieve over a ROW-SHARDED corpus (one process per GPU) -- the `ids_only` flow of
`retrieve_evidence` (app/retrieve.py:392-573) when the tables do not fit, or are not wanted, on one GPU.

SPMD: every rank calls :func:`sharded_retrieve_ids` with the same request and gets the same response.
Each rank holds, per table, a `DenseStore` over its contiguous row range (ids stay global), a tech-token
index over those rows and a `ShardedSearcher`.  Per request and table:

  filter      evaluated locally (K6); COUNT(*) = sum of the local counts (one small all-reduce)
  dense lane  local K1 / K2 scan, then the exchange + merge of `ShardedSearcher` (K4p peer-memory kernel
              over NVLink, or NCCL all-gather + K4): every rank ends with the global top-k
  tech lane   local top-`limit` in the lane's order (call_started_at DESC, id ASC); the (started_at, id)
              pairs of all ranks travel with the local counts in ONE small object all-gather per table and
              are merged by the same key
  rows        the ids_only response needs ids and scores only; payload columns stay on the owning rank
              (`_owned_rows`)
  fusion      the bit-exact RRF kernel (K5) and the ids_only combine, on every rank (inputs are identical)

The result equals `retrieve_ids` over the unsharded corpus (`tests/test_gpu_sharded.py`).
"""
from __future__ import annotations

from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np

from . import retrieve as R
from . import _ffi
from ._ffi import DenseEngineError
from .dist import ShardedSearcher
from .embeddings import EmbeddingClientError, embed_texts, embeddings_enabled
from .lexical import extract_tech_tokens


class ShardedEngine:
    """One rank's shard of the engine: the local `DenseEngine` (stores + tech indexes over this rank's rows)
    plus one `ShardedSearcher` per table.  Construction is collective (the searchers set up their exchange)."""

    def __init__(self, local: R.DenseEngine, group=None, transport: str = "auto"):
        import torch.distributed as dist
        self.local = local
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        devices = {store.device for store in local.stores.values()}
        if len(devices) != 1:
            raise DenseEngineError("a rank's shard stores must live on one device")
        R.settings.cadence_gpu_device = devices.pop()        # the RRF kernel runs on this rank's GPU
        self.searchers: Dict[str, ShardedSearcher] = {
            table: ShardedSearcher(store, group, transport=transport, max_nq=64, max_k=64)
            for table, store in sorted(local.stores.items())}

    def close(self) -> None:
        for s in self.searchers.values():
            s.close()


def _gather_objects(obj: Any, eng: ShardedEngine) -> List[Any]:
    import torch.distributed as dist
    if eng.world == 1:
        return [obj]
    out: List[Any] = [None] * eng.world
    dist.all_gather_object(out, obj, group=eng.group)
    return out


def _owned_rows(store, ids: Sequence[int]) -> Dict[int, Dict[str, Any]]:
    """Row dicts (id column, call_id, payload) of the ids that live on this rank."""
    cols = store.host_columns()
    own = cols["ids"]
    out: Dict[int, Dict[str, Any]] = {}
    if own.size == 0:
        return out
    want = np.asarray(list(ids), dtype=np.int64)
    pos = np.searchsorted(own, want)
    for i, p in zip(want.tolist(), pos.tolist()):
        if p < own.size and int(own[p]) == i:
            slot = int(cols["call_slot"][p])
            call_id = store.call_ids_by_slot[slot] if slot < len(store.call_ids_by_slot) else slot
            row = {store.key_field: i, "call_id": call_id}
            row.update(store.payload.get(i, {}))
            out[i] = row
    return out


def _sharded_table(eng: ShardedEngine, conn, table: str, q32: Optional[np.ndarray], tech_tokens: Sequence[str],
                   filters, call_ids, dense_limit: int, tech_limit: int) -> Tuple[List[Dict[str, Any]], List[Dict[str, Any]], int]:
    """(tech rows, dense rows, COUNT(*)) of one table, identical on every rank.  Rows carry what the ids_only
    response needs -- the id, and the score on the dense lane; payload columns stay on the owning rank
    (`_owned_rows` fetches them when a caller wants them).  Host traffic per table: ONE small object all-gather
    (the tech lane's (call_started_at, id) candidates + the local COUNT(*)); the dense lane's exchange runs on the
    devices (`ShardedSearcher`)."""
    import torch
    import torch.distributed as dist
    store = conn.store(table)
    searcher = eng.searchers[table]
    key = store.key_field
    dense = q32 is not None
    if dense:
        want = max(1, int(R.settings.embeddings_dim))
        if q32.shape[0] != want or q32.shape[0] != store.dim:
            raise DenseEngineError(f"expected {want} dimensions, not {q32.shape[0]}", _ffi.CDR_ERR_UNSUPPORTED)
    # ---- local halves: tech-lane winners with their sort key, filter bitmap + local COUNT(*)
    local_tech = R._fetch_tech(conn, table, tech_tokens, filters, call_ids, tech_limit)
    cols = store.host_columns()
    keyed = []
    for row in local_tech:
        p = int(np.searchsorted(cols["ids"], row[key]))
        keyed.append((-int(cols["started_at"][p]), int(row[key])))
    allow, local_count = R._filter_bitmap(conn, table, filters, call_ids) if dense else (None, 0)
    # ---- one exchange of host objects: merge the tech lane by its ORDER BY key, sum the counts
    parts = _gather_objects((keyed, int(local_count)), eng)
    merged = sorted(pair for part, _c in parts for pair in part)
    tech_rows = [{key: item[1]} for item in merged[:tech_limit]]
    count = sum(int(c) for _p, c in parts)
    # ---- dense lane: local scan, device-side exchange + merge
    dense_rows: List[Dict[str, Any]] = []
    if dense and count > 0:
        on_gpu = eng.world == 1 or dist.get_backend(eng.group) == "nccl"      # gloo: the CPU test tier
        dev = f"cuda:{store.device}" if on_gpu else "cpu"
        qd = torch.from_numpy(q32[None, :]).to(dev)
        ids, scores, n = searcher.search(qd, dense_limit, allow, mode="exact")
        m = int(n[0].item())
        dense_rows = [{key: i, "score": sc} for i, sc in zip(ids[0, :m].tolist(), scores[0, :m].tolist())]
    return tech_rows, dense_rows, count if dense else 0


def sharded_retrieve_ids(eng: ShardedEngine, query: str, filters: Optional[R.RetrieveFilters] = None,
                         bm25_chunks: Sequence[Mapping[str, Any]] = (),
                         bm25_artifacts: Sequence[Mapping[str, Any]] = (), debug: bool = False) -> Dict[str, Any]:
    """`retrieve_ids` over the sharded corpus.  Collective: call it on every rank with the same arguments."""
    import torch.distributed as dist
    query = query.strip()
    if not query:
        return {"retrieved_ids": []}
    tech_tokens = extract_tech_tokens(query)
    dense_enabled = embeddings_enabled()
    dense_error: Optional[str] = None
    dense_model_id: Optional[str] = None
    q32: Optional[np.ndarray] = None
    if dense_enabled:
        # rank 0 talks to the embedding service; everybody searches with the same vector
        box: List[Any] = [None]
        if eng.rank == 0:
            try:
                embedded = embed_texts([query])
                box[0] = ("ok", embedded.model, R._embedding_f32(embedded.vectors[0]))
            except EmbeddingClientError as exc:
                box[0] = ("error", str(exc), None)
        if eng.world > 1:
            dist.broadcast_object_list(box, src=0, group=eng.group)
        if box[0][0] == "ok":
            dense_model_id, q32 = box[0][1], box[0][2]
        else:
            dense_enabled, dense_error = False, box[0][1]

    modes: Dict[str, Optional[str]] = {"chunks": None, "artifact_chunks": None}
    candidates = {"chunks": 0, "artifact_chunks": 0}
    lanes: Dict[str, Tuple[List[Dict[str, Any]], List[Dict[str, Any]]]] = {"chunks": ([], []), "artifact_chunks": ([], [])}
    limits = {"chunks": R.DEFAULT_DENSE_CHUNK_TOPK, "artifact_chunks": R.DEFAULT_DENSE_ARTIFACT_CHUNK_TOPK}
    with eng.local.connect() as conn:
        call_ids = R._resolve_call_ids(conn, filters)
        for attempt in (0, 1):
            try:
                for table in ("chunks", "artifact_chunks"):
                    if table not in eng.local.stores:
                        continue
                    tech, dense, count = _sharded_table(eng, conn, table, q32 if dense_enabled else None, tech_tokens,
                                                        filters, call_ids, limits[table], R.DEFAULT_TECH_TOPK)
                    lanes[table] = (tech, dense)
                    if dense_enabled:
                        candidates[table] = count
                        modes[table] = R._choose_dense_mode(count, filters, call_ids)
                break
            except DenseEngineError as exc:       # deterministic on every rank: fail open to lexical-only
                if not dense_enabled or attempt == 1 or not R._dense_failure_is_recoverable(exc):
                    raise
                dense_enabled, dense_error = False, str(exc)
                modes = {"chunks": None, "artifact_chunks": None}
                candidates = {"chunks": 0, "artifact_chunks": 0}
    (tech_chunks, dense_chunks), (tech_artifacts, dense_artifacts) = lanes["chunks"], lanes["artifact_chunks"]
    chunk_lanes: Dict[str, Sequence[Mapping[str, Any]]] = {"bm25": list(bm25_chunks), "tech_tokens": tech_chunks}
    artifact_lanes: Dict[str, Sequence[Mapping[str, Any]]] = {"bm25": list(bm25_artifacts), "tech_tokens": tech_artifacts}
    if dense_enabled:
        chunk_lanes["dense"] = dense_chunks
        artifact_lanes["dense"] = dense_artifacts
    chunk_ranked = R._rrf_merge(chunk_lanes, "chunk_id")
    artifact_ranked = R._rrf_merge(artifact_lanes, "artifact_chunk_id")
    return R._ids_response(chunk_ranked, artifact_ranked, bm25_chunks, bm25_artifacts, tech_chunks, tech_artifacts,
                           dense_chunks, dense_artifacts, dense_enabled, dense_model_id, dense_error, modes,
                           candidates, debug)
