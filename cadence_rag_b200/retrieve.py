"""Dense-lane retrieval facade -- the reference's app/retrieve.py call surface over the GPU store.

Same function names, argument order and return shapes as the reference for the functions on the
hot path; ``conn`` is a :class:`DenseConnection` instead of a SQLAlchemy
connection and ``query_embedding`` may be the reference's ``"[...]"`` literal, a float sequence
or a tensor:

  _resolve_call_ids            app/retrieve.py:46-90
  _estimate_dense_candidates   app/retrieve.py:303-323   exact COUNT(*) -> K6 popcount
  _dense_has_scoping           app/retrieve.py:267-274
  _choose_dense_mode           app/retrieve.py:277-287
  _configure_dense_session     app/retrieve.py:290-300
  _fetch_chunks_dense          app/retrieve.py:326-354   SQL ORDER BY <=> LIMIT -> K1 / K2
  _fetch_artifacts_dense       app/retrieve.py:357-389
  _fetch_chunks_tech / _fetch_artifacts_tech   app/retrieve.py:183-242 (host inverted index)
  _rrf_merge                   app/retrieve.py:245-260   -> K5 (bit-exact)
  _vector_literal              app/retrieve.py:263-264
  _build_debug_lane            app/retrieve.py:35-43
  retrieve_ids                 the ids_only branch of retrieve_evidence, app/retrieve.py:392-573
"""
from __future__ import annotations

from array import array
from dataclasses import dataclass
from datetime import datetime
from typing import Any, Dict, Iterable, List, Mapping, Optional, Sequence, Set, Tuple

import numpy as np

from . import _ffi
from ._ffi import DenseEngineError
from .config import settings
from .embeddings import EmbeddingClientError, embed_texts, embeddings_enabled
from .lexical import DeviceTechIndex, TechTokenIndex, extract_tech_tokens
from .store import DenseStore

DEFAULT_RRF_K = 60
DEFAULT_CHUNK_BM25_TOPK = 50
DEFAULT_ARTIFACT_CHUNK_BM25_TOPK = 10
DEFAULT_DENSE_CHUNK_TOPK = 50
DEFAULT_DENSE_ARTIFACT_CHUNK_TOPK = 10
DEFAULT_TECH_TOPK = 50


@dataclass
class RetrieveFilters:
    """Duck-type of the reference's pydantic model (app/schemas.py:77-83)."""
    date_from: Optional[datetime] = None
    date_to: Optional[datetime] = None
    call_ids: Optional[List[Any]] = None
    external_id: Optional[str] = None
    external_source: Optional[str] = None
    call_tags: Optional[List[str]] = None


class DenseConnection:
    """Stands where ``with engine.connect() as conn`` stands in the reference
    (app/retrieve.py:445): request-scoped access to the resident stores.  Filter bitmaps built
    for a request are cached on the connection, so the COUNT(*) estimate and the search share one
    K6 launch."""

    def __init__(self, engine: "DenseEngine"):
        self.engine = engine
        self._bitmaps: Dict[Tuple, Tuple[Any, int]] = {}
        self.session: Dict[str, Any] = {}       # what _configure_dense_session last set (SET LOCAL ... in the reference)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self._bitmaps.clear()
        return False

    def store(self, table_name: str) -> DenseStore:
        try:
            return self.engine.stores[table_name]
        except KeyError:
            raise DenseEngineError(f"no resident store for table {table_name!r}") from None


class DenseEngine:
    """Holds the resident stores ("chunks", "artifact_chunks"), their tech-token indexes and the
    external_id -> call_id map used by filters."""

    def __init__(self):
        self.stores: Dict[str, DenseStore] = {}
        self.tech_indexes: Dict[str, TechTokenIndex] = {}
        self.device_tech_indexes: Dict[str, DeviceTechIndex] = {}
        self.external_ids: Dict[Tuple[str, Optional[str]], Set[Any]] = {}
        # optional provider of the opaque BM25 lanes for retrieve_evidence(payload):
        # (query, filters) -> (chunk rows, artifact rows), ranked, each row carrying its id field
        self.bm25_lanes = None
        import threading
        self._tech_lock = threading.Lock()
        self._retired_tech_indexes: List[DeviceTechIndex] = []

    def register(self, store: DenseStore, tech_index: Optional[TechTokenIndex] = None,
                 device_tech_lane: bool = True) -> None:
        """``tech_index``: host inverted index of the table's tech_tokens.  With ``device_tech_lane``
        (default) it is also uploaded as a :class:`DeviceTechIndex` and the lane runs on the GPU."""
        self.stores[store.table_name] = store
        if tech_index is not None:
            self.tech_indexes[store.table_name] = tech_index
            if device_tech_lane:
                self.device_tech_indexes[store.table_name] = DeviceTechIndex(tech_index, store)

    def device_tech_index(self, table_name: str) -> Optional[DeviceTechIndex]:
        """The table's device tech-token index, rebuilt first when the store grew or the host index changed since it
        was built (a sealed store keeps serving while it grows; the lane must see the new rows like the reference's
        SQL does).  None when the table has no device lane."""
        dev = self.device_tech_indexes.get(table_name)
        if dev is not None and dev.stale():
            with self._tech_lock:
                dev = self.device_tech_indexes.get(table_name)
                if dev is not None and dev.stale():
                    fresh = DeviceTechIndex(dev.host_index, dev.store)
                    self.device_tech_indexes[table_name] = fresh
                    self._retired_tech_indexes.append(dev)      # requests in flight may still hold it: freed at close()
                    dev = fresh
        return dev

    def close(self) -> None:
        for dev in list(self.device_tech_indexes.values()) + self._retired_tech_indexes:
            dev.close()
        self.device_tech_indexes.clear()
        self._retired_tech_indexes.clear()

    def register_call(self, call_id, external_id: Optional[str] = None, external_source: Optional[str] = None) -> None:
        if external_id is not None:
            self.external_ids.setdefault((external_id, external_source), set()).add(call_id)

    def connect(self) -> DenseConnection:
        return DenseConnection(self)


# --------------------------------------------------------------------------- dense-lane failures
_RECOVERABLE_DENSE_CODES = (_ffi.CDR_ERR_UNSUPPORTED, _ffi.CDR_ERR_OOM, _ffi.CDR_ERR_NO_DEVICE)


def _dense_failure_is_recoverable(exc: DenseEngineError) -> bool:
    """The reference fails open to lexical-only on `EmbeddingClientError` alone (app/retrieve.py:426-432).  The engine's
    counterpart: a request the dense lane cannot serve (unsupported shape, out of device memory, no device) degrades
    the same way; a CUDA error, a bad argument or a store in the wrong state is an engine fault -- it is logged and
    re-raised, never hidden behind a lexical-only answer."""
    ok = getattr(exc, "code", _ffi.CDR_ERR_INVALID) in _RECOVERABLE_DENSE_CODES
    if not ok:
        import logging
        logging.getLogger("cadence_rag_b200").error("dense lane failed (code %s): %s", getattr(exc, "code", None), exc)
    return ok


# --------------------------------------------------------------------------- small pure helpers
def _build_debug_lane(rows: Sequence[Mapping[str, Any]], id_field: str) -> List[Dict[str, Any]]:
    return [{id_field: row[id_field], "rank": rank, "score": row.get("score")}
            for rank, row in enumerate(rows, start=1)]


def _vector_literal(values: Sequence[float]) -> str:
    return "[" + ",".join(format(float(v), ".10g") for v in values) + "]"


def _query_vector(query_embedding) -> np.ndarray:
    """Reference literal / float sequence / tensor -> float32 vector.  The text form is parsed the
    way pgvector's vector_in does (strtof per element => exact float32 rounding)."""
    if isinstance(query_embedding, str):
        body = query_embedding.strip()
        if not (body.startswith("[") and body.endswith("]")):
            raise DenseEngineError("malformed vector literal")
        return np.array([np.float32(tok) for tok in body[1:-1].split(",")], dtype=np.float32)
    if hasattr(query_embedding, "detach"):
        return query_embedding.detach().to("cpu").numpy().astype(np.float32).reshape(-1)
    if isinstance(query_embedding, (list, tuple)):
        # the embedding client's float list: through a C double array (half the time of numpy's per-element conversion of
        # a list; the double -> float32 rounding is the same round-to-nearest)
        try:
            return np.frombuffer(array("d", query_embedding), dtype=np.float64).astype(np.float32)
        except (TypeError, OverflowError):
            pass
    return np.asarray(query_embedding, dtype=np.float32).reshape(-1)


def _resolve_call_ids(conn: DenseConnection, filters: Optional[RetrieveFilters]) -> Optional[List[Any]]:
    if not filters:
        return None
    call_ids: Optional[Set[Any]] = set(filters.call_ids) if filters.call_ids else None
    if filters.external_id:
        resolved: Set[Any] = set()
        for (ext_id, ext_src), ids in conn.engine.external_ids.items():
            if ext_id != filters.external_id:
                continue
            # reference: no source given => any source; else IS NOT DISTINCT FROM :external_source
            if filters.external_source is None or ext_src == filters.external_source:
                resolved |= ids
        if call_ids:
            call_ids &= resolved
        else:
            call_ids = resolved
    if call_ids is None:
        return None
    return sorted(call_ids, key=str)


def _dense_has_scoping(filters: Optional[RetrieveFilters], call_ids: Optional[Sequence[Any]]) -> bool:
    if call_ids is not None:
        return True
    if not filters:
        return False
    return bool(filters.date_from or filters.date_to or filters.call_tags)


def _choose_dense_mode(estimated_rows: int, filters: Optional[RetrieveFilters],
                       call_ids: Optional[Sequence[Any]]) -> str:
    if estimated_rows <= 0:
        return "exact"
    if _dense_has_scoping(filters, call_ids):
        if estimated_rows <= max(settings.embeddings_exact_scan_threshold, 0):
            return "exact"
    return "ann"


def _configure_dense_session(conn: "DenseConnection", mode: str) -> None:
    """app/retrieve.py:290-300.  The reference steers the Postgres planner per statement: mode "exact" turns
    index scans off (forced sequential scan), mode "ann" turns them on and sets hnsw.iterative_scan /
    hnsw.ef_search.  Here the mode selects the lane -- "exact": the fp32 HBM-bound scan; "ann": the batched
    bf16 tensor-core lane is PERMITTED (a single query is still answered exactly whenever fp32 rows are
    resident) -- and the same settings are recorded on the connection for the debug / notes payload."""
    if mode == "ann":
        conn.session = {"mode": "ann", "enable_indexscan": "on", "enable_bitmapscan": "on",
                        "hnsw.iterative_scan": "relaxed_order",
                        "hnsw.ef_search": max(1, int(settings.embeddings_hnsw_ef_search))}
        return
    conn.session = {"mode": "exact", "enable_indexscan": "off", "enable_bitmapscan": "off"}


# --------------------------------------------------------------------------- filters -> bitmap
def _filter_spec(store: DenseStore, filters: Optional[RetrieveFilters],
                 call_ids: Optional[Sequence[Any]]) -> Dict[str, Any]:
    """_build_filter_clause (app/retrieve.py:93-120) as store codes: call slots, dates, tag mask.  Like the
    reference, no filters object means no predicate at all (call_ids are only ever resolved FROM filters)."""
    if not filters:
        return dict(call_slots=None, date_from=None, date_to=None, tag_mask=None)
    date_from = (filters.date_from or None) if filters else None
    date_to = (filters.date_to or None) if filters else None
    tags = list(filters.call_tags) if (filters and filters.call_tags) else None
    slots = None
    if call_ids is not None:
        slots = []
        for c in call_ids:
            s = store.slot_of_call(c)
            if s is None and store.synthetic is not None and isinstance(c, (int, np.integer)):
                s = int(c)          # synthetic stores: call id == slot number
            if s is not None:
                slots.append(s)
    tag_mask = None
    if tags is not None:
        # tags beyond the 64 device bits live host-side as call-slot sets (store.tag_filter): such a filter becomes
        # a call-slot set, intersected with the scoped calls when there are any
        tag_mask, tag_slots = store.tag_filter(tags)
        if tag_slots is not None:
            slots = tag_slots if slots is None else sorted(set(slots) & set(tag_slots))
    return dict(call_slots=slots, date_from=date_from, date_to=date_to, tag_mask=tag_mask)


def _filter_bitmap(conn: DenseConnection, table_name: str, filters: Optional[RetrieveFilters],
                   call_ids: Optional[Sequence[Any]]):
    """(allow tensor or None, candidate count) for WHERE <filters> AND embedding IS NOT NULL."""
    store = conn.store(table_name)
    spec = _filter_spec(store, filters, call_ids)
    key = (table_name, None if spec["call_slots"] is None else tuple(spec["call_slots"]),
           spec["date_from"], spec["date_to"], spec["tag_mask"])
    hit = conn._bitmaps.get(key)
    if hit is not None:
        return hit
    if all(v is None for v in spec.values()):
        result = (None, int(store.info()["n_valid"]))
    else:
        result = store.filter_bitmap(**spec)
    conn._bitmaps[key] = result
    return result


def _estimate_dense_candidates(conn: DenseConnection, table_name: str,
                               filters: Optional[RetrieveFilters],
                               call_ids: Optional[Sequence[Any]]) -> int:
    return int(_filter_bitmap(conn, table_name, filters, call_ids)[1])


# --------------------------------------------------------------------------- dense lanes
def _rows_from_hits(store: DenseStore, ids: np.ndarray, scores: np.ndarray) -> List[Dict[str, Any]]:
    out: List[Dict[str, Any]] = []
    if len(ids) == 0:
        return out
    cols = store.host_columns()
    pos = np.searchsorted(cols["ids"], ids)
    for i, sc, p in zip(ids.tolist(), scores.tolist(), pos.tolist()):
        slot = int(cols["call_slot"][p])
        call_id = store.call_ids_by_slot[slot] if slot < len(store.call_ids_by_slot) else slot
        row = {store.key_field: i, "call_id": call_id}
        row.update(store.payload.get(i, {}))
        row["score"] = sc
        out.append(row)
    return out


def _fetch_dense(conn: DenseConnection, table_name: str, query_embedding, filters, call_ids,
                 mode: str, limit: int) -> List[Dict[str, Any]]:
    _configure_dense_session(conn, mode)
    store = conn.store(table_name)
    if limit <= 0:
        return []
    q = _query_vector(query_embedding)
    want = max(1, int(settings.embeddings_dim))
    if q.shape[0] != want or q.shape[0] != store.dim:
        raise DenseEngineError(f"expected {want} dimensions, not {q.shape[0]}", _ffi.CDR_ERR_UNSUPPORTED)
    allow, count = _filter_bitmap(conn, table_name, filters, call_ids)
    if count <= 0:
        return []
    # mode "ann" permits the approximate lane (reference: HNSW ef_search, app/retrieve.py:291-298);
    # a single query is answered exactly by the HBM-bound scan whenever fp32 rows are resident.
    use_batch = (mode == "ann" and store.has_bf16 and
                 (not store.has_fp32 or settings.cadence_gpu_ann_min_batch <= 1))
    use_scan = (mode == "ann" and store.has_bf16 and not use_batch and int(settings.cadence_gpu_ann_bf16_scan)
                and not _dense_has_scoping(filters, call_ids) and store.dim in (256, 512, 768, 1024))
    if use_batch:
        ids, scores, n = store.search_batch(q, limit, allow)
    elif use_scan:
        ids, scores, n = store.search_scan_bf16(q, limit, allow)
    else:
        ids, scores, n = store.search_exact(q, limit, allow)
    m = int(n[0])
    return _rows_from_hits(store, ids[0, :m], scores[0, :m])


def _fetch_chunks_dense(conn: DenseConnection, query_embedding, filters: Optional[RetrieveFilters],
                        call_ids: Optional[Sequence[Any]], mode: str, limit: int) -> List[Dict[str, Any]]:
    return _fetch_dense(conn, "chunks", query_embedding, filters, call_ids, mode, limit)


def _fetch_artifacts_dense(conn: DenseConnection, query_embedding, filters: Optional[RetrieveFilters],
                           call_ids: Optional[Sequence[Any]], mode: str, limit: int) -> List[Dict[str, Any]]:
    return _fetch_dense(conn, "artifact_chunks", query_embedding, filters, call_ids, mode, limit)


def fetch_dense_batch(conn: DenseConnection, table_name: str, queries, filters, call_ids, mode: str,
                      limit: int):
    """Batched form (nq queries at once; not in the reference, which embeds one query per
    request): returns (ids[nq,limit], scores[nq,limit], n[nq]).  mode "ann" with at least
    ``cadence_gpu_ann_min_batch`` queries runs on the tcgen05 lane."""
    store = conn.store(table_name)
    allow, count = _filter_bitmap(conn, table_name, filters, call_ids)
    nq = int(queries.shape[0])
    if mode == "ann" and store.has_bf16 and (nq >= settings.cadence_gpu_ann_min_batch or not store.has_fp32):
        if not store.has_fp32 or _batch_lane_is_faster(store, nq, count):
            return store.search_batch(queries, limit, allow)
    if (mode == "ann" and store.has_bf16 and nq <= 8 and int(settings.cadence_gpu_ann_bf16_scan)
            and not _dense_has_scoping(filters, call_ids) and store.dim in (256, 512, 768, 1024)):
        return store.search_scan_bf16(queries, limit, allow)          # a few unscoped queries share passes over the bf16 rows
    return store.search_exact(queries, limit, allow, shared=True)


def _batch_lane_is_faster(store: DenseStore, nq: int, candidate_rows: int) -> bool:
    """Lane choice inside mode "ann" for a batch.  The bf16 tensor-core lane multiplies every 256-row tile that
    keeps at least one row (call-level filters keep runs of rows: ~2.5 x the candidate rows are touched) at
    ~1.2 PFLOP/s, with ~1.5 ms of small-segment overhead under a filter.  The exact lane with shared reads streams
    only the candidate rows when the filter keeps <= rows/2 (gather launch), once per 8 queries (shared-memory
    bound, ~4.3 TB/s of tile reads) or once per 3 queries for batches of <= 6.  Measured rates from
    profiles/r01/README.md; small candidate sets always stay on the exact lane."""
    rows, dim = store.rows, store.dim
    scanned = candidate_rows if candidate_rows * 2 <= rows else rows
    if scanned * dim * 4 <= (1 << 26):
        return False
    touched = rows if candidate_rows * 2.5 >= rows else int(candidate_rows * 2.5)
    nq_pad = -(-nq // 128) * 128
    t_batch = max(2.0 * nq_pad * touched * dim / 1.2e15, touched * dim * 2 / 6.7e12) + (2e-4 if touched == rows else 1.5e-3)
    if nq > 6:
        passes, rate = -(-nq // 8), 4.3e12
    else:
        passes, rate = -(-nq // 3), (5.4e12 if scanned < rows else 6.5e12)
    t_exact = passes * (scanned * dim * 4 / rate + 2e-5) + 5e-5
    return t_batch < t_exact


# --------------------------------------------------------------------------- tech_tokens lane
def _fetch_tech(conn: DenseConnection, table_name: str, tokens: Sequence[str], filters, call_ids,
                limit: int) -> List[Dict[str, Any]]:
    if not tokens:
        return []
    index = conn.engine.tech_indexes.get(table_name)
    if index is None:
        return []
    store = conn.store(table_name)
    # the tech lane's WHERE has no `embedding IS NOT NULL` term (app/retrieve.py:195-208), so the
    # predicate is evaluated on the host columns of the posting rows, not through the K6 bitmap
    spec = _filter_spec(store, filters, call_ids)
    dev_index = conn.engine.device_tech_index(table_name)
    if dev_index is not None and not dev_index.fits(tokens):
        dev_index = None                           # more tokens than the kernel's table: the host index serves
    cols = store.host_columns()
    if dev_index is not None:                      # GPU lane (f-1); ids come back already ordered
        hit_ids = np.asarray(dev_index.query_ids(tokens, limit, **spec), dtype=np.int64)
        rows = np.searchsorted(cols["ids"], hit_ids)
    else:
        rows = index.query(tokens, cols, limit, **spec)
    cols = store.host_columns()
    out = []
    for p in rows.tolist():
        i = int(cols["ids"][p])
        slot = int(cols["call_slot"][p])
        call_id = store.call_ids_by_slot[slot] if slot < len(store.call_ids_by_slot) else slot
        row = {store.key_field: i, "call_id": call_id}
        row.update(store.payload.get(i, {}))
        out.append(row)
    return out


def _fetch_chunks_tech(conn, tokens, filters, call_ids, limit):
    return _fetch_tech(conn, "chunks", tokens, filters, call_ids, limit)


def _fetch_artifacts_tech(conn, tokens, filters, call_ids, limit):
    return _fetch_tech(conn, "artifact_chunks", tokens, filters, call_ids, limit)


# --------------------------------------------------------------------------- RRF (K5)
def _rrf_merge(lanes: Mapping[str, Sequence[Mapping[str, Any]]], key_field: str,
               k: int = DEFAULT_RRF_K) -> List[Tuple[Dict[str, Any], Set[str], float]]:
    """Bit-exact GPU restatement of app/retrieve.py:245-260.  Keys must be integer ids
    (chunk_id / artifact_chunk_id are BIGSERIAL)."""
    names = list(lanes.keys())
    if not names:
        return []
    items: Dict[int, Mapping[str, Any]] = {}
    flat: List[int] = []
    offsets = [0]
    for name in names:
        for row in lanes[name]:
            key = int(row[key_field])
            items.setdefault(key, row)
            flat.append(key)
        offsets.append(len(flat))
    if not flat:
        return []
    if len(flat) > _ffi.CDR_RRF_MAX_ITEMS:
        raise DenseEngineError(f"rrf: {len(flat)} lane items exceed {_ffi.CDR_RRF_MAX_ITEMS}")
    ids, scores, masks, n = rrf_merge_batch(np.asarray(flat, dtype=np.int64),
                                            np.asarray(offsets, dtype=np.int32), 1, len(names), k,
                                            max_out=len(flat))
    out = []
    for i in range(int(n[0])):
        key = int(ids[0, i])
        hit = {names[l] for l in range(len(names)) if (int(masks[0, i]) >> l) & 1}
        out.append((items[key], hit, float(scores[0, i])))
    return out


def rrf_merge_batch(lane_ids: np.ndarray, lane_offsets: np.ndarray, nq: int, n_lanes: int, k: int,
                    max_out: int):
    """Host-buffer K5 call for nq queries: see cdr_rrf_merge_host in include/cadence_dense.h."""
    import torch
    _ffi.require_device()
    lane_ids = np.ascontiguousarray(lane_ids, dtype=np.int64)
    lane_offsets = np.ascontiguousarray(lane_offsets, dtype=np.int32)
    assert lane_offsets.shape[0] == nq * n_lanes + 1
    out_ids = np.empty((nq, max_out), dtype=np.int64)
    out_sc = np.empty((nq, max_out), dtype=np.float64)
    out_mask = np.empty((nq, max_out), dtype=np.uint32)
    out_n = np.empty((nq,), dtype=np.int32)
    with torch.cuda.device(settings.cadence_gpu_device):
        _ffi.check(_ffi.lib().cdr_rrf_merge_host(_ffi.ptr(lane_ids), _ffi.ptr(lane_offsets), nq, n_lanes, k,
                                                 max_out, _ffi.ptr(out_ids), _ffi.ptr(out_sc),
                                                 _ffi.ptr(out_mask), _ffi.ptr(out_n), _ffi.stream_ptr()),
                   "cdr_rrf_merge_host")
    return out_ids, out_sc, out_mask, out_n


# --------------------------------------------------------------------------- fused per-table request
_LANE_NAMES = ("bm25", "tech_tokens", "dense")


def _embedding_f32(values: Sequence[float]) -> np.ndarray:
    """What `CAST(:q AS vector(1024))` yields for `_vector_literal(values)`, without the text round
    trip when it is provably the identity: a float32-representable value printed with 10 significant
    digits (`.10g`, 9 suffice for float32) parses back to itself.  Anything else (fp64 decimals from a
    JSON body, NaN) takes the literal path."""
    a64 = np.asarray(values, dtype=np.float64).reshape(-1)
    a32 = a64.astype(np.float32)
    if np.array_equal(a32.astype(np.float64), a64):
        return a32
    return _query_vector(_vector_literal(values))


def _embeddings_f32(vectors: Sequence[Sequence[float]]) -> np.ndarray:
    """`_embedding_f32` for a batch: one conversion for all rows; a row that is not float32-representable
    takes the literal path on its own."""
    try:
        a64 = np.asarray(vectors, dtype=np.float64)
    except ValueError:                                   # ragged rows: let the per-row path report them
        return np.stack([_embedding_f32(v) for v in vectors])
    if a64.ndim != 2:
        return np.stack([_embedding_f32(v) for v in vectors])
    a32 = a64.astype(np.float32)
    exact = (a32.astype(np.float64) == a64).all(axis=1)
    for i in np.flatnonzero(~exact).tolist():
        a32[i] = _query_vector(_vector_literal(vectors[i]))
    return a32


def _fused_path_ok(engine: DenseEngine, table: str, dense: bool, token_lists: Sequence[Sequence[str]] = ()) -> bool:
    """The fused C call serves a table when its dense lane is the exact fp32 scan and its tech lane
    (if any) is device resident and can take every request's tokens; other configurations take the
    step-by-step path."""
    store = engine.stores[table]
    if table in engine.tech_indexes and table not in engine.device_tech_indexes:
        return False
    dev = engine.device_tech_indexes.get(table)
    if dev is not None and any(len(t) > dev.MAX_TOKENS and not dev.fits(t) for t in token_lists):
        return False
    if dense and not store.has_fp32:
        return False
    if dense and store.has_bf16 and settings.cadence_gpu_ann_min_batch <= 1:
        return False        # single queries are routed to the batched bf16 lane by configuration
    return True


def _group_dense_lane(store: DenseStore, filters, call_ids, n_requests: int) -> int:
    """Dense lane of a group of requests inside the fused call.  Unscoped requests always plan "ann"
    (app/retrieve.py:277-287: the exact mode needs scoping), where the reference walks its HNSW index; a group of at
    least `cadence_gpu_ann_min_batch` of them goes to the batched bf16 tensor-core lane when that is the faster one.
    Scoped groups stay on the exact scan: their mode depends on COUNT(*), which is computed inside the same call."""
    if not store.has_bf16 or _dense_has_scoping(filters, call_ids):
        return _ffi.CDR_DENSE_LANE_EXACT_F32
    # Without the tensor-core lane: one request scans the bf16 rows (0.30 ms per 1 M rows; the fp32 rows 0.58 ms), a few
    # share passes over the bf16 rows in pairs (two 0.39 ms, eight 1.05 ms; the shared fp32 scan 0.62 / 1.36 ms); from
    # about sixteen on the register-tiled fp32 scan (16 queries per pass) is the faster one (profiles/r02/README.md 10)
    scan = (_ffi.CDR_DENSE_LANE_SCAN_BF16
            if int(settings.cadence_gpu_ann_bf16_scan) and store.dim in (256, 512, 768, 1024) and n_requests <= 8
            else _ffi.CDR_DENSE_LANE_EXACT_F32)
    if n_requests < max(2, int(settings.cadence_gpu_ann_min_batch)):
        return scan
    if _batch_lane_is_faster(store, n_requests, store.rows):
        return _ffi.CDR_DENSE_LANE_BATCH_BF16
    return scan


def _hybrid_table_batch(conn: DenseConnection, table: str, q32: Optional[np.ndarray],
                        token_lists: Sequence[Sequence[str]], filters, call_ids,
                        bm25_rows: Sequence[Sequence[Mapping[str, Any]]],
                        dense_limit: int, tech_limit: int = DEFAULT_TECH_TOPK,
                        rrf_k: int = DEFAULT_RRF_K, per_request_filters: bool = False,
                        ids_only: bool = False) -> List[Dict[str, Any]]:
    """All lanes of one table + their fusion for nq requests through ONE fused C call.  q32: [nq, dim] float32
    or None (dense lane disabled); token_lists and bm25_rows: one entry per request.  The requests share
    `filters` / `call_ids`, or -- with per_request_filters -- each request brings its own (sequences of length
    nq): requests are then grouped by filter and every group runs its own filter / lane launches inside the
    same call (`cdr_hybrid_retrieve_groups_host`).  Returns, per request, {"tech": rows, "dense": rows,
    "count": COUNT(*), "ranked": [(row, lane-name set, score)]} with the rows / order the step-by-step
    functions produce (lane rows carry the id -- and the score on the dense lane -- only).  With `ids_only` the
    per-row dicts are skipped: each entry is {"count", "fused": [(id, score)] in fused order} (the response without
    a debug payload needs nothing else)."""
    store = conn.store(table)
    key = store.key_field
    nq = len(token_lists)
    if q32 is not None:
        want = max(1, int(settings.embeddings_dim))
        if q32.shape[1] != want or q32.shape[1] != store.dim:
            raise DenseEngineError(f"expected {want} dimensions, not {q32.shape[1]}", _ffi.CDR_ERR_UNSUPPORTED)
    order = list(range(nq))                      # request -> position in the C call
    group_specs = group_off = None
    if per_request_filters:
        specs = [_filter_spec(store, filters[i], call_ids[i]) for i in range(nq)]
        groups: Dict[Tuple, List[int]] = {}
        for i, sp in enumerate(specs):
            gk = (None if sp["call_slots"] is None else tuple(sp["call_slots"]), sp["date_from"], sp["date_to"], sp["tag_mask"])
            groups.setdefault(gk, []).append(i)
        order, group_specs, group_off = [], [], [0]
        for members in groups.values():
            order += members
            lane = _group_dense_lane(store, filters[members[0]], call_ids[members[0]], len(members)) if q32 is not None else 0
            group_specs.append(dict(specs[members[0]], dense_lane=lane) if lane else specs[members[0]])
            group_off.append(len(order))
        spec = None
    else:
        spec = _filter_spec(store, filters, call_ids)
        lane = _group_dense_lane(store, filters, call_ids, nq) if q32 is not None else 0
        if lane:
            spec = dict(spec, dense_lane=lane)
    token_lists = [token_lists[i] for i in order]
    bm25_rows = [bm25_rows[i] for i in order]
    if q32 is not None and per_request_filters:
        q32 = np.ascontiguousarray(q32[order])
    dev_index = conn.engine.device_tech_index(table) if any(token_lists) else None
    tok = nt = None
    if dev_index is not None:
        tok, nt = dev_index.encode_tokens([list(t) for t in token_lists])
    counts = [len(rows) for rows in bm25_rows]
    bm25_ids = np.fromiter((int(r[key]) for rows in bm25_rows for r in rows), dtype=np.int64, count=sum(counts))
    bm25_off = np.zeros(nq + 1, dtype=np.int32)
    np.cumsum(counts, out=bm25_off[1:])
    res = store.hybrid_retrieve(q32, dense_limit, tech_index=dev_index, token_ids=tok, n_tokens=nt,
                                tech_limit=tech_limit, bm25_ids=bm25_ids, bm25_offsets=bm25_off, rrf_k=rrf_k,
                                filter_spec=spec, filter_specs=group_specs, group_offsets=group_off)
    if per_request_filters:                       # COUNT(*) of the group each position belongs to
        pos_count = [0] * nq
        for gi in range(len(group_specs)):
            for pos in range(group_off[gi], group_off[gi + 1]):
                pos_count[pos] = res["count"][gi]
    else:
        pos_count = [res["count"]] * nq
    n_lanes = 3 if q32 is not None else 2
    # one bulk conversion per output array: indexing numpy scalars row by row costs more than the C call
    fused_ids, fused_scores, fused_n = res["fused_ids"].tolist(), res["fused_scores"].tolist(), res["fused_n"].tolist()
    if not ids_only:
        tech_ids, tech_n, fused_mask = res["tech_ids"].tolist(), res["tech_n"].tolist(), res["fused_mask"].tolist()
        hit_sets = [frozenset(_LANE_NAMES[l] for l in range(n_lanes) if (m >> l) & 1) for m in range(1 << n_lanes)]
    if q32 is not None and not ids_only:
        dense_ids, dense_scores, dense_n = res["dense_ids"].tolist(), res["dense_scores"].tolist(), res["dense_n"].tolist()
    out_pos = []
    for qi in range(nq if not ids_only else 0):
        # the ids_only response needs ids, ranks and scores only: the SELECT-list columns (call_id, payload) are
        # looked up by retrieve_evidence for the few rows that make it into the pack, not for every lane row
        tech_rows = [{key: i} for i in tech_ids[qi][:tech_n[qi]]]
        dense_rows: List[Dict[str, Any]] = []
        if q32 is not None:
            m = dense_n[qi]
            dense_rows = [{key: i, "score": sc} for i, sc in zip(dense_ids[qi][:m], dense_scores[qi][:m])]
        items: Dict[int, Mapping[str, Any]] = {}
        for row in bm25_rows[qi]:
            items.setdefault(int(row[key]), row)
        for row in tech_rows:
            items.setdefault(row[key], row)
        for row in dense_rows:
            items.setdefault(row[key], row)
        n_f = fused_n[qi]
        ranked = [(items[i], set(hit_sets[mk]), sc)
                  for i, mk, sc in zip(fused_ids[qi][:n_f], fused_mask[qi][:n_f], fused_scores[qi][:n_f])]
        out_pos.append({"tech": tech_rows, "dense": dense_rows, "count": pos_count[qi] if q32 is not None else 0,
                        "ranked": ranked})
    if ids_only:
        out_pos = [{"count": pos_count[qi] if q32 is not None else 0,
                    "fused": list(zip(fused_ids[qi][:fused_n[qi]], fused_scores[qi][:fused_n[qi]]))} for qi in range(nq)]
    out: List[Dict[str, Any]] = [None] * nq      # back to request order
    for pos, i in enumerate(order):
        out[i] = out_pos[pos]
    return out


def _hybrid_table(conn: DenseConnection, table: str, q32: Optional[np.ndarray], tech_tokens: Sequence[str],
                  filters: Optional[RetrieveFilters], call_ids: Optional[Sequence[Any]],
                  bm25_rows: Sequence[Mapping[str, Any]], dense_limit: int,
                  tech_limit: int = DEFAULT_TECH_TOPK, rrf_k: int = DEFAULT_RRF_K) -> Dict[str, Any]:
    """One request (see _hybrid_table_batch)."""
    return _hybrid_table_batch(conn, table, None if q32 is None else q32[None, :], [list(tech_tokens)], filters,
                               call_ids, [list(bm25_rows)], dense_limit, tech_limit, rrf_k)[0]


def retrieve_ids_batch(engine: DenseEngine, queries: Sequence[str], filters=None,
                       bm25_chunks: Optional[Sequence[Sequence[Mapping[str, Any]]]] = None,
                       bm25_artifacts: Optional[Sequence[Sequence[Mapping[str, Any]]]] = None,
                       debug: bool = False) -> List[Dict[str, Any]]:
    """`retrieve_ids` for a batch of concurrent requests (not in the reference, which serves one query per
    request): one embedding call, then ONE fused C call per table for the whole batch instead of one per
    request.  `filters`: one RetrieveFilters / None shared by all requests, or a list with one entry per
    request (requests with equal filters form a group and share corpus reads on the dense lane).  Returns one
    response per query, each equal to what `retrieve_ids` returns for it.  Falls back to per-request calls when a
    table cannot take the fused path."""
    n = len(queries)
    per_request = isinstance(filters, (list, tuple))
    if per_request and len(filters) != n:
        raise ValueError("one filter per query is required")
    filter_of = (lambda i: filters[i]) if per_request else (lambda i: filters)
    bm25_chunks = [list(r) for r in bm25_chunks] if bm25_chunks is not None else [[] for _ in range(n)]
    bm25_artifacts = [list(r) for r in bm25_artifacts] if bm25_artifacts is not None else [[] for _ in range(n)]
    if len(bm25_chunks) != n or len(bm25_artifacts) != n:
        raise ValueError("one BM25 lane per query is required")
    cleaned = [q.strip() for q in queries]
    live = [i for i, q in enumerate(cleaned) if q]
    tables = [t for t in ("chunks", "artifact_chunks") if t in engine.stores]
    dense_enabled = embeddings_enabled()
    live_tokens = [extract_tech_tokens(cleaned[i]) for i in live]
    if not live or not tables or not all(_fused_path_ok(engine, t, dense_enabled, live_tokens) for t in tables):
        return [retrieve_ids(engine, q, filter_of(i), bm25_chunks=bm25_chunks[i], bm25_artifacts=bm25_artifacts[i], debug=debug)
                for i, q in enumerate(queries)]
    dense_error: Optional[str] = None
    dense_model_id: Optional[str] = None
    q32 = None
    lean = not debug            # no debug payload: ids and fused scores straight from the call's output arrays
    if dense_enabled:
        try:
            embedded = embed_texts([cleaned[i] for i in live])
            dense_model_id = embedded.model
            q32 = _embeddings_f32(embedded.vectors)
        except EmbeddingClientError as exc:
            dense_enabled = False
            dense_error = str(exc)
    token_lists = live_tokens
    limits = {"chunks": DEFAULT_DENSE_CHUNK_TOPK, "artifact_chunks": DEFAULT_DENSE_ARTIFACT_CHUNK_TOPK}
    bm25 = {"chunks": [bm25_chunks[i] for i in live], "artifact_chunks": [bm25_artifacts[i] for i in live]}
    with engine.connect() as conn:
        if per_request:
            live_filters = [filters[i] for i in live]
            call_ids = [_resolve_call_ids(conn, f) for f in live_filters]
        else:
            live_filters = filters
            call_ids = _resolve_call_ids(conn, filters)
        per_table: Dict[str, List[Dict[str, Any]]] = {}
        try:
            for t in tables:
                per_table[t] = _hybrid_table_batch(conn, t, q32 if dense_enabled else None, token_lists, live_filters,
                                                   call_ids, bm25[t], limits[t], per_request_filters=per_request,
                                                   ids_only=lean)
        except DenseEngineError as exc:       # fail open to lexical-only, like EmbeddingClientError
            if not dense_enabled or not _dense_failure_is_recoverable(exc):
                raise
            dense_enabled = False
            dense_error = str(exc)
            for t in tables:
                per_table[t] = _hybrid_table_batch(conn, t, None, token_lists, live_filters, call_ids, bm25[t], limits[t],
                                                   per_request_filters=per_request, ids_only=lean)
    empty = {"tech": [], "dense": [], "count": 0, "ranked": []}
    responses: List[Dict[str, Any]] = [{"retrieved_ids": []} for _ in range(n)]
    if lean:
        kinds = ("artifact_chunk", "chunk")
        for j, i in enumerate(live):
            combined: List[Tuple[float, int, int]] = []
            for kind, t, key, extra in ((0, "artifact_chunks", "artifact_chunk_id", bm25_artifacts[i]),
                                        (1, "chunks", "chunk_id", bm25_chunks[i])):
                if t in per_table:
                    combined += [(-sc, kind, item) for item, sc in per_table[t][j]["fused"]]
                else:                                    # table not resident: its BM25 lane alone
                    combined += [(-sc, kind, row[key]) for row, _l, sc in _rrf_merge({"bm25": extra, "tech_tokens": []}, key)]
            combined.sort()
            responses[i] = {"retrieved_ids": [f"{kinds[kind]}:{item}" for _, kind, item in combined]}
        return responses
    for j, i in enumerate(live):
        ch = per_table["chunks"][j] if "chunks" in per_table else empty
        ar = per_table["artifact_chunks"][j] if "artifact_chunks" in per_table else empty
        modes: Dict[str, Optional[str]] = {"chunks": None, "artifact_chunks": None}
        candidates = {"chunks": 0, "artifact_chunks": 0}
        if dense_enabled:
            for t in tables:
                candidates[t] = per_table[t][j]["count"]
                modes[t] = _choose_dense_mode(candidates[t], filter_of(i), call_ids[j] if per_request else call_ids)
        chunk_ranked, artifact_ranked = ch["ranked"], ar["ranked"]
        if "chunks" not in per_table:
            chunk_ranked = _rrf_merge({"bm25": bm25_chunks[i], "tech_tokens": []}, "chunk_id")
        if "artifact_chunks" not in per_table:
            artifact_ranked = _rrf_merge({"bm25": bm25_artifacts[i], "tech_tokens": []}, "artifact_chunk_id")
        responses[i] = _ids_response(chunk_ranked, artifact_ranked, bm25_chunks[i], bm25_artifacts[i], ch["tech"], ar["tech"],
                                     ch["dense"], ar["dense"], dense_enabled, dense_model_id, dense_error, modes,
                                     candidates, debug)
    return responses


class RequestBatcher:
    """Micro-batcher in front of concurrent /retrieve requests.  The reference serves every request on its own
    threadpool thread (app/main.py:184-186); here client threads hand their request to one worker, which drains
    the queue -- up to `max_batch` requests, waiting at most `max_wait_s` for company -- and serves the whole batch
    with ONE `retrieve_ids_batch` call: one embedding call, one fused C call per table (requests with equal filters
    share corpus reads), one synchronisation.  Each client gets exactly the response `retrieve_ids` would give it."""

    def __init__(self, engine: DenseEngine, max_batch: int = 64, max_wait_s: float = 2e-4, workers: int = 1):
        import queue
        import threading
        self.engine = engine
        self.max_batch = max(1, int(max_batch))
        self.max_wait_s = float(max_wait_s)
        self._queue: "queue.Queue" = queue.Queue()
        self._closed = False
        self._stats = threading.Lock()
        self.batches_served = 0
        self.requests_served = 0
        # one worker by default: a second one (it would build response dicts while the first waits for the GPU in the
        # GIL-free C call) halves the batch size and measured no better (profiles/r01/README.md)
        self._workers = [threading.Thread(target=self._run, name=f"cadence-request-batcher-{i}", daemon=True)
                         for i in range(max(1, int(workers)))]
        for w in self._workers:
            w.start()

    def retrieve_ids(self, query: str, filters: Optional[RetrieveFilters] = None,
                     bm25_chunks: Sequence[Mapping[str, Any]] = (), bm25_artifacts: Sequence[Mapping[str, Any]] = (),
                     debug: bool = False) -> Dict[str, Any]:
        import threading
        if self._closed:
            raise DenseEngineError("request batcher is closed", _ffi.CDR_ERR_STATE)
        ticket = {"args": (query, filters, list(bm25_chunks), list(bm25_artifacts), debug), "done": threading.Event(),
                  "result": None, "error": None}
        self._queue.put(ticket)
        while not ticket["done"].wait(timeout=1.0):
            if not any(w.is_alive() for w in self._workers):
                raise DenseEngineError("request batcher has no live worker", _ffi.CDR_ERR_STATE)
        if ticket["error"] is not None:
            raise ticket["error"]
        return ticket["result"]

    def close(self) -> None:
        self._closed = True
        for _ in self._workers:
            self._queue.put(None)
        for w in self._workers:
            w.join(timeout=5)

    def _run(self) -> None:
        import queue
        import time
        import torch
        try:
            stores = list(self.engine.stores.values())
            if stores and torch.cuda.is_available():
                with torch.cuda.stream(torch.cuda.Stream(device=stores[0].device)):   # each worker on its own stream
                    self._serve(queue, time)
            else:
                self._serve(queue, time)
        except BaseException as exc:   # noqa: BLE001 - a dead worker must not leave clients waiting
            self._closed = True
            while True:
                try:
                    t = self._queue.get_nowait()
                except queue.Empty:
                    break
                if t is not None:
                    t["error"] = DenseEngineError(f"request batcher worker died: {exc!r}", _ffi.CDR_ERR_STATE)
                    t["done"].set()

    def _serve(self, queue, time) -> None:
        while True:
            first = self._queue.get()
            if first is None:
                return
            batch = [first]
            deadline = time.perf_counter() + self.max_wait_s
            stop = False
            while len(batch) < self.max_batch:
                remaining = deadline - time.perf_counter()
                try:
                    nxt = self._queue.get(timeout=remaining) if remaining > 0 else self._queue.get_nowait()
                except queue.Empty:
                    break
                if nxt is None:
                    stop = True
                    break
                batch.append(nxt)
            try:
                want_debug = any(t["args"][4] for t in batch)
                responses = retrieve_ids_batch(self.engine, [t["args"][0] for t in batch], [t["args"][1] for t in batch],
                                               bm25_chunks=[t["args"][2] for t in batch],
                                               bm25_artifacts=[t["args"][3] for t in batch], debug=want_debug)
                for t, resp in zip(batch, responses):
                    if want_debug and not t["args"][4]:
                        resp = {k: v for k, v in resp.items() if k != "debug"}
                    t["result"] = resp
            except BaseException:   # noqa: BLE001
                # fault isolation: one client's malformed request (filters, BM25 rows) or one engine fault must not fail
                # up to max_batch unrelated clients -- the members are retried one at a time and each gets its own
                # response or its own error
                for t in batch:
                    try:
                        q, f, bc, ba, dbg = t["args"]
                        t["result"] = retrieve_ids(self.engine, q, f, bm25_chunks=bc, bm25_artifacts=ba, debug=dbg)
                    except BaseException as exc:   # noqa: BLE001 - handed to the client that sent it
                        t["error"] = exc
            with self._stats:
                self.batches_served += 1
                self.requests_served += len(batch)
            for t in batch:
                t["done"].set()
            if stop:
                return


# --------------------------------------------------------------------------- ids_only retrieve
def retrieve_ids(engine: DenseEngine, query: str, filters: Optional[RetrieveFilters] = None,
                 bm25_chunks: Sequence[Mapping[str, Any]] = (),
                 bm25_artifacts: Sequence[Mapping[str, Any]] = (), debug: bool = False) -> Dict[str, Any]:
    """The ``return_style == "ids_only"`` flow of retrieve_evidence (app/retrieve.py:392-573) over
    the GPU engine.  The BM25 lane is out of scope (pg_search); its ranked rows are accepted as an
    opaque input lane, exactly where the reference feeds them into _rrf_merge."""
    query = query.strip()
    if not query:
        return {"retrieved_ids": []}
    if not debug:
        # without a debug payload the response is built straight from the fused call's output arrays
        resident = [t for t in ("chunks", "artifact_chunks") if t in engine.stores]
        if resident and all(_fused_path_ok(engine, t, embeddings_enabled(), [extract_tech_tokens(query)]) for t in resident):
            return retrieve_ids_batch(engine, [query], filters, bm25_chunks=[list(bm25_chunks)],
                                      bm25_artifacts=[list(bm25_artifacts)])[0]
    tech_tokens = extract_tech_tokens(query)
    dense_enabled = embeddings_enabled()
    dense_error: Optional[str] = None
    dense_model_id: Optional[str] = None
    query_embedding = None       # float32 vector == CAST(_vector_literal(v) AS vector(D)) (see _embedding_f32)
    if dense_enabled:
        try:
            embedded = embed_texts([query])
            dense_model_id = embedded.model
            query_embedding = _embedding_f32(embedded.vectors[0])
        except EmbeddingClientError as exc:
            dense_enabled = False
            dense_error = str(exc)

    tech_chunks: List[Dict[str, Any]] = []
    tech_artifacts: List[Dict[str, Any]] = []
    dense_chunks: List[Dict[str, Any]] = []
    dense_artifacts: List[Dict[str, Any]] = []
    modes: Dict[str, Optional[str]] = {"chunks": None, "artifact_chunks": None}
    candidates = {"chunks": 0, "artifact_chunks": 0}
    tables = [t for t in ("chunks", "artifact_chunks") if t in engine.stores]
    fused = bool(tables) and all(_fused_path_ok(engine, t, dense_enabled, [tech_tokens]) for t in tables)
    if fused:
        # one C call per table: filter + lanes + RRF on the device, one sync (csrc/hybrid.cu)
        limits = {"chunks": DEFAULT_DENSE_CHUNK_TOPK, "artifact_chunks": DEFAULT_DENSE_ARTIFACT_CHUNK_TOPK}
        bm25 = {"chunks": list(bm25_chunks), "artifact_chunks": list(bm25_artifacts)}
        with engine.connect() as conn:
            call_ids = _resolve_call_ids(conn, filters)
            per_table: Dict[str, Dict[str, Any]] = {}
            try:
                for t in tables:
                    per_table[t] = _hybrid_table(conn, t, query_embedding if dense_enabled else None, tech_tokens,
                                                 filters, call_ids, bm25[t], limits[t])
            except DenseEngineError as exc:       # fail open to lexical-only, like EmbeddingClientError
                if not dense_enabled or not _dense_failure_is_recoverable(exc):
                    raise
                dense_enabled = False
                dense_error = str(exc)
                for t in tables:
                    per_table[t] = _hybrid_table(conn, t, None, tech_tokens, filters, call_ids, bm25[t], limits[t])
        empty = {"tech": [], "dense": [], "count": 0, "ranked": []}
        ch, ar = per_table.get("chunks", empty), per_table.get("artifact_chunks", empty)
        tech_chunks, tech_artifacts = ch["tech"], ar["tech"]
        dense_chunks, dense_artifacts = ch["dense"], ar["dense"]
        if dense_enabled:
            for t in tables:
                candidates[t] = per_table[t]["count"]
                modes[t] = _choose_dense_mode(candidates[t], filters, call_ids)
        chunk_ranked, artifact_ranked = ch["ranked"], ar["ranked"]
        if "chunks" not in per_table:
            chunk_ranked = _rrf_merge({"bm25": list(bm25_chunks), "tech_tokens": []}, "chunk_id")
        if "artifact_chunks" not in per_table:
            artifact_ranked = _rrf_merge({"bm25": list(bm25_artifacts), "tech_tokens": []}, "artifact_chunk_id")
        return _ids_response(chunk_ranked, artifact_ranked, bm25_chunks, bm25_artifacts, tech_chunks, tech_artifacts,
                             dense_chunks, dense_artifacts, dense_enabled, dense_model_id, dense_error, modes,
                             candidates, debug)
    with engine.connect() as conn:
        call_ids = _resolve_call_ids(conn, filters)
        if "chunks" in engine.stores:
            tech_chunks = _fetch_chunks_tech(conn, tech_tokens, filters, call_ids, DEFAULT_TECH_TOPK)
        if "artifact_chunks" in engine.stores:
            tech_artifacts = _fetch_artifacts_tech(conn, tech_tokens, filters, call_ids, DEFAULT_TECH_TOPK)
        if dense_enabled and query_embedding is not None:
            try:
                for table, fetch, topk, sink in (
                        ("chunks", _fetch_chunks_dense, DEFAULT_DENSE_CHUNK_TOPK, dense_chunks),
                        ("artifact_chunks", _fetch_artifacts_dense, DEFAULT_DENSE_ARTIFACT_CHUNK_TOPK, dense_artifacts)):
                    if table not in engine.stores:
                        continue
                    candidates[table] = _estimate_dense_candidates(conn, table, filters, call_ids)
                    modes[table] = _choose_dense_mode(candidates[table], filters, call_ids)
                    sink.extend(fetch(conn, query_embedding, filters, call_ids, modes[table], topk))
            except DenseEngineError as exc:   # fail open to lexical-only, like EmbeddingClientError
                if not _dense_failure_is_recoverable(exc):
                    raise
                dense_enabled = False
                dense_error = str(exc)
                dense_chunks.clear()
                dense_artifacts.clear()

    chunk_lanes: Dict[str, Sequence[Mapping[str, Any]]] = {"bm25": list(bm25_chunks), "tech_tokens": tech_chunks}
    artifact_lanes: Dict[str, Sequence[Mapping[str, Any]]] = {"bm25": list(bm25_artifacts), "tech_tokens": tech_artifacts}
    if dense_enabled:
        chunk_lanes["dense"] = dense_chunks
        artifact_lanes["dense"] = dense_artifacts
    chunk_ranked = _rrf_merge(chunk_lanes, "chunk_id")
    artifact_ranked = _rrf_merge(artifact_lanes, "artifact_chunk_id")
    return _ids_response(chunk_ranked, artifact_ranked, bm25_chunks, bm25_artifacts, tech_chunks, tech_artifacts,
                         dense_chunks, dense_artifacts, dense_enabled, dense_model_id, dense_error, modes,
                         candidates, debug)


def _ids_response(chunk_ranked, artifact_ranked, bm25_chunks, bm25_artifacts, tech_chunks, tech_artifacts,
                  dense_chunks, dense_artifacts, dense_enabled, dense_model_id, dense_error, modes, candidates,
                  debug: bool) -> Dict[str, Any]:
    """ids_only combine (app/retrieve.py:552-573) + the debug payload."""
    # sort key (-score, kind order {artifact_chunk: 0, chunk: 1}, id) carried as the tuple itself
    combined = [(-score, 0, row["artifact_chunk_id"]) for row, _lanes, score in artifact_ranked]
    combined += [(-score, 1, row["chunk_id"]) for row, _lanes, score in chunk_ranked]
    combined.sort()
    kinds = ("artifact_chunk", "chunk")
    response: Dict[str, Any] = {"retrieved_ids": [f"{kinds[kind]}:{item_id}" for _, kind, item_id in combined]}
    if debug:
        chunk_dbg = {"bm25": _build_debug_lane(list(bm25_chunks), "chunk_id"),
                     "tech_tokens": _build_debug_lane(tech_chunks, "chunk_id")}
        art_dbg = {"bm25": _build_debug_lane(list(bm25_artifacts), "artifact_chunk_id"),
                   "tech_tokens": _build_debug_lane(tech_artifacts, "artifact_chunk_id")}
        if dense_enabled:
            chunk_dbg["dense"] = _build_debug_lane(dense_chunks, "chunk_id")
            art_dbg["dense"] = _build_debug_lane(dense_artifacts, "artifact_chunk_id")
        response["debug"] = {
            "lanes": {"chunks": chunk_dbg, "artifacts": art_dbg},
            "limits": {"bm25_chunk_topk": DEFAULT_CHUNK_BM25_TOPK,
                       "bm25_artifact_chunk_topk": DEFAULT_ARTIFACT_CHUNK_BM25_TOPK,
                       "tech_token_topk": DEFAULT_TECH_TOPK,
                       "dense_chunk_topk": DEFAULT_DENSE_CHUNK_TOPK if dense_enabled else 0,
                       "dense_artifact_chunk_topk": DEFAULT_DENSE_ARTIFACT_CHUNK_TOPK if dense_enabled else 0},
            "dense": {"enabled": dense_enabled, "model_id": dense_model_id, "error": dense_error,
                      "modes": dict(modes), "candidate_rows": dict(candidates)},
            "fused": {"chunks": [(r["chunk_id"], sorted(l), s) for r, l, s in chunk_ranked],
                      "artifacts": [(r["artifact_chunk_id"], sorted(l), s) for r, l, s in artifact_ranked]},
        }
    return response


# --------------------------------------------------------------------------- evidence pack (f-3)
DEFAULT_MAX_ARTIFACTS = 2
DEFAULT_MAX_QUOTES_PER_CALL = 2
DEFAULT_SNIPPET_CHARS = 800


@dataclass
class Budget:
    """app/schemas.py:71-73."""
    max_evidence_items: int = 8
    max_total_chars: int = 6000


def _clip(text: str, max_chars: int) -> str:
    """app/retrieve.py:27-32: hard cut with a trailing ellipsis inside the budget."""
    if max_chars <= 0:
        return ""
    return text if len(text) <= max_chars else text[: max_chars - 1].rstrip() + "…"


_default_engine: Optional[DenseEngine] = None


def set_default_engine(engine: Optional[DenseEngine]) -> None:
    """The engine ``retrieve_evidence(payload)`` serves from -- the counterpart of the reference's module-level
    SQLAlchemy ``engine`` (app/db.py:11, used at app/retrieve.py:445)."""
    global _default_engine
    _default_engine = engine


def default_engine() -> DenseEngine:
    if _default_engine is None:
        raise DenseEngineError("retrieve_evidence(payload): no engine installed (call set_default_engine(engine) once at "
                               "start-up, as the reference creates its SQLAlchemy engine at import time)", _ffi.CDR_ERR_STATE)
    return _default_engine


def _budget_dict(budget) -> Dict[str, int]:
    """``budget.model_dump()`` of the reference (app/retrieve.py:413) for its pydantic Budget and for ours."""
    return {"max_evidence_items": int(budget.max_evidence_items), "max_total_chars": int(budget.max_total_chars)}


def retrieve_evidence(engine, query: Optional[str] = None, filters: Optional[RetrieveFilters] = None,
                      budget: Optional[Budget] = None, intent: str = "auto",
                      return_style: str = "evidence_pack_json", debug: bool = False,
                      bm25_chunks: Sequence[Mapping[str, Any]] = (),
                      bm25_artifacts: Sequence[Mapping[str, Any]] = ()) -> Dict[str, Any]:
    """The reference's /retrieve response contract (app/retrieve.py:392-688) over the GPU engine:
    lanes -> RRF -> either ``ids_only`` or the budgeted evidence pack (<= 2 artifacts, <= 2 quotes
    per call, snippet <= 800 chars, total chars <= budget) with the same ``notes.retrieval`` block.
    Payload columns (`content`, `artifact_id`, `kind`, `text`, `speaker`, `start_ts_ms`,
    `end_ts_ms`) come from the rows registered with the stores.

    Two call shapes:
      ``retrieve_evidence(payload)`` -- the reference's own signature (app/retrieve.py:392): ``payload`` is its
        ``RetrieveRequest`` (app/schemas.py:86-93; any object with query / intent / filters / budget / return_style /
        debug), served from the engine installed with :func:`set_default_engine`; the opaque BM25 lanes come from
        the engine's ``bm25_lanes`` hook when one is registered (pg_search is out of scope);
      ``retrieve_evidence(engine, query, filters, budget, ...)`` -- the explicit form (tests, replay)."""
    if not isinstance(engine, DenseEngine) and hasattr(engine, "query") and hasattr(engine, "return_style"):
        payload, eng = engine, default_engine()
        lanes = ((), ())
        if getattr(eng, "bm25_lanes", None) is not None:
            lanes = eng.bm25_lanes(payload.query.strip(), payload.filters)
        return retrieve_evidence(eng, payload.query, payload.filters, payload.budget, payload.intent,
                                 payload.return_style, payload.debug, lanes[0], lanes[1])
    from uuid import uuid4
    query_id = str(uuid4())
    budget = budget or Budget()
    query = query.strip()                                   # app/retrieve.py:394
    if not query:
        if return_style == "ids_only":
            return {"query_id": query_id, "retrieved_ids": []}
        return {"query_id": query_id, "intent": intent, "budget": _budget_dict(budget), "artifacts": [],
                "quotes": [], "notes": {"error": "empty query"}}
    inner = retrieve_ids(engine, query, filters, bm25_chunks=bm25_chunks, bm25_artifacts=bm25_artifacts, debug=True)
    dbg = inner["debug"]
    if return_style == "ids_only":
        out = {"query_id": query_id, "retrieved_ids": inner["retrieved_ids"]}
        if debug:
            out["debug"] = dbg
        return out

    rows_by_id = {"chunks": {}, "artifacts": {}}
    for kind, table, key in (("chunks", "chunks", "chunk_id"), ("artifacts", "artifact_chunks", "artifact_chunk_id")):
        store = engine.stores.get(table)
        for ident, lanes, score in dbg["fused"][kind]:
            row = {key: ident}
            if store is not None:
                row.update(store.payload.get(ident, {}))
                cols = store.host_columns()
                pos = int(np.searchsorted(cols["ids"], ident))
                if pos < len(cols["ids"]) and cols["ids"][pos] == ident:
                    slot = int(cols["call_slot"][pos])
                    row["call_id"] = store.call_ids_by_slot[slot] if slot < len(store.call_ids_by_slot) else slot
            rows_by_id[kind][ident] = (row, lanes)

    remaining = budget.max_total_chars
    used = 0
    artifacts_out: List[Dict[str, Any]] = []
    for ident, lanes, _s in dbg["fused"]["artifacts"]:
        if used >= budget.max_evidence_items or len(artifacts_out) >= min(DEFAULT_MAX_ARTIFACTS, budget.max_evidence_items):
            break
        if remaining <= 0:
            break
        row, _ = rows_by_id["artifacts"][ident]
        snippet = _clip(str(row.get("content", "")), min(DEFAULT_SNIPPET_CHARS, remaining))
        remaining -= len(snippet)
        artifacts_out.append({"evidence_id": f"A-{ident}", "call_id": str(row.get("call_id")),
                              "artifact_id": row.get("artifact_id"), "artifact_chunk_id": ident,
                              "kind": row.get("kind"), "snippet": snippet, "why_relevant": " + ".join(sorted(lanes))})
        used += 1
    quotes_out: List[Dict[str, Any]] = []
    per_call: Dict[str, int] = {}
    for ident, lanes, _s in dbg["fused"]["chunks"]:
        if used >= budget.max_evidence_items or remaining <= 0:
            break
        row, _ = rows_by_id["chunks"][ident]
        call_id = str(row.get("call_id"))
        if per_call.get(call_id, 0) >= DEFAULT_MAX_QUOTES_PER_CALL:
            continue
        snippet = _clip(str(row.get("text", "")), min(DEFAULT_SNIPPET_CHARS, remaining))
        remaining -= len(snippet)
        quotes_out.append({"evidence_id": f"Q-{ident}", "call_id": call_id, "chunk_id": ident,
                           "speaker": row.get("speaker"), "start_ts_ms": row.get("start_ts_ms"),
                           "end_ts_ms": row.get("end_ts_ms"), "snippet": snippet,
                           "why_relevant": " + ".join(sorted(lanes))})
        per_call[call_id] = per_call.get(call_id, 0) + 1
        used += 1

    dense = dbg["dense"]
    modes = dense["modes"]
    planner = ("lexical_only" if not dense["enabled"] else
               ("ann" if "ann" in (modes.get("chunks"), modes.get("artifact_chunks")) else "exact"))
    response = {
        "query_id": query_id, "intent": intent, "budget": _budget_dict(budget),
        "artifacts": artifacts_out, "quotes": quotes_out,
        "notes": {"retrieval": {
            "planner": planner,
            "dense_topk": max(DEFAULT_DENSE_CHUNK_TOPK, DEFAULT_DENSE_ARTIFACT_CHUNK_TOPK) if dense["enabled"] else 0,
            "lex_topk": DEFAULT_CHUNK_BM25_TOPK, "artifact_chunk_lex_topk": DEFAULT_ARTIFACT_CHUNK_BM25_TOPK,
            "reranked_from": None, "bm25_chunk_topk": DEFAULT_CHUNK_BM25_TOPK,
            "bm25_artifact_chunk_topk": DEFAULT_ARTIFACT_CHUNK_BM25_TOPK, "tech_token_topk": DEFAULT_TECH_TOPK,
            "tech_tokens": extract_tech_tokens(query.strip()),
            "lanes": {"bm25": True, "tech_tokens": True, "dense": dense["enabled"]},
            "dense_model_id": dense["model_id"], "dense_error": dense["error"],
            "dense_modes": {"chunks": modes.get("chunks"), "artifact_chunks": modes.get("artifact_chunks")},
            "dense_candidate_rows": dict(dense["candidate_rows"]),
            "hnsw_ef_search": settings.embeddings_hnsw_ef_search if dense["enabled"] else None,
        }},
    }
    if debug:
        response["debug"] = dbg
    return response


# --------------------------------------------------------------------------- hierarchical scoping (f-4)
def fetch_chunks_dense_hierarchical(conn: DenseConnection, query_embedding, filters: Optional[RetrieveFilters],
                                    call_ids: Optional[Sequence[Any]], artifact_limit: int = DEFAULT_DENSE_ARTIFACT_CHUNK_TOPK,
                                    chunk_limit: int = DEFAULT_DENSE_CHUNK_TOPK, max_calls: int = 8) -> Dict[str, Any]:
    """The two-stage dense retrieval the reference specifies but only half implements
    (APP_SPEC.md:621-638; today both corpora are searched globally, app/retrieve.py:472-487):
      1. dense search over ``artifact_chunks`` (summaries / decisions / action items),
      2. the calls of the best artifact hits become a shortlist (first-seen order, <= max_calls),
      3. ``chunks`` is searched *scoped to those calls* -- the K6 bitmap restricts the K1 scan, and the
         planner sees a scoped query, so small shortlists run in mode "exact".
    Returns the artifact rows, the shortlist, the scoped chunk rows and the planner decisions."""
    filters = filters or RetrieveFilters()        # the shortlist below scopes the chunk search even without user filters
    art_count = _estimate_dense_candidates(conn, "artifact_chunks", filters, call_ids)
    art_mode = _choose_dense_mode(art_count, filters, call_ids)
    artifacts = _fetch_artifacts_dense(conn, query_embedding, filters, call_ids, art_mode, artifact_limit)
    shortlist: List[Any] = []
    for row in artifacts:
        if row["call_id"] not in shortlist:
            shortlist.append(row["call_id"])
        if len(shortlist) >= max_calls:
            break
    if call_ids is not None:
        allowed = set(call_ids)
        shortlist = [c for c in shortlist if c in allowed]
    chunk_count = _estimate_dense_candidates(conn, "chunks", filters, shortlist)
    chunk_mode = _choose_dense_mode(chunk_count, filters, shortlist)
    chunks = _fetch_chunks_dense(conn, query_embedding, filters, shortlist, chunk_mode, chunk_limit)
    return {"artifacts": artifacts, "call_shortlist": shortlist, "chunks": chunks,
            "modes": {"artifact_chunks": art_mode, "chunks": chunk_mode},
            "candidate_rows": {"artifact_chunks": art_count, "chunks": chunk_count}}
