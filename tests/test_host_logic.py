"""CPU tier: host-side mirror of the reference interface (no GPU compute).

Reads like the reference's own unit tests (tests/unit/test_retrieve_planner.py,
tests/unit/test_embeddings_client.py, tests/unit/test_ingest_utils.py) pointed at this package,
plus the C-ABI load/export check.
"""
import ctypes
import json
import os
import re
import subprocess
import sys
from datetime import datetime, timezone
from uuid import uuid4

import numpy as np
import pytest

import cadence_rag_b200 as pkg
from cadence_rag_b200 import _ffi, embeddings, lexical, retrieve
from cadence_rag_b200.config import settings
from cadence_rag_b200.retrieve import RetrieveFilters, _choose_dense_mode
from oracle import ports

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pure(golden_dir):
    with open(os.path.join(golden_dir, "reference_pure.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module", autouse=True)
def built_library():
    if not os.path.exists(_ffi.library_path()):
        import __graft_entry__
        __graft_entry__.build()


# ---------------------------------------------------------------- planner (reference tests mirrored)
def test_choose_dense_mode_exact_for_small_scoped_sets(monkeypatch):
    monkeypatch.setattr(settings, "embeddings_exact_scan_threshold", 2000)
    filters = RetrieveFilters(call_ids=[uuid4()])
    assert _choose_dense_mode(estimated_rows=200, filters=filters, call_ids=filters.call_ids) == "exact"


def test_choose_dense_mode_ann_for_large_scoped_sets(monkeypatch):
    monkeypatch.setattr(settings, "embeddings_exact_scan_threshold", 2000)
    filters = RetrieveFilters(date_from=datetime.now(timezone.utc), date_to=datetime.now(timezone.utc))
    assert _choose_dense_mode(estimated_rows=5000, filters=filters, call_ids=None) == "ann"


def test_choose_dense_mode_ann_for_unscoped_queries(monkeypatch):
    monkeypatch.setattr(settings, "embeddings_exact_scan_threshold", 5000)
    assert _choose_dense_mode(estimated_rows=100, filters=None, call_ids=None) == "ann"


def test_choose_dense_mode_exact_when_no_candidates(monkeypatch):
    monkeypatch.setattr(settings, "embeddings_exact_scan_threshold", 10)
    assert _choose_dense_mode(estimated_rows=0, filters=None, call_ids=None) == "exact"


def test_configure_dense_session_records_reference_settings(monkeypatch):
    """app/retrieve.py:290-300: "ann" -> index scans on, relaxed_order, ef_search = max(1, int(setting));
    anything else -> index scans off."""
    from cadence_rag_b200.retrieve import DenseConnection, DenseEngine, _configure_dense_session
    conn = DenseConnection(DenseEngine())
    monkeypatch.setattr(settings, "embeddings_hnsw_ef_search", 80)
    _configure_dense_session(conn, "ann")
    assert conn.session == {"mode": "ann", "enable_indexscan": "on", "enable_bitmapscan": "on",
                            "hnsw.iterative_scan": "relaxed_order", "hnsw.ef_search": 80}
    monkeypatch.setattr(settings, "embeddings_hnsw_ef_search", 0)
    _configure_dense_session(conn, "ann")
    assert conn.session["hnsw.ef_search"] == 1
    _configure_dense_session(conn, "exact")
    assert conn.session == {"mode": "exact", "enable_indexscan": "off", "enable_bitmapscan": "off"}


def test_planner_matches_reference_golden_table(pure, monkeypatch):
    now = datetime(2026, 2, 9, tzinfo=timezone.utc)
    for case in pure["planner"]:
        monkeypatch.setattr(settings, "embeddings_exact_scan_threshold", case["threshold"])
        kind = case["kind"]
        filters, call_ids = None, None
        if kind == "nofilter_callids":
            call_ids = ["c1"]
        elif kind == "empty_callids":
            call_ids = []
        elif kind == "date_from":
            filters = RetrieveFilters(date_from=now)
        elif kind == "date_to":
            filters = RetrieveFilters(date_to=now)
        elif kind == "tags":
            filters = RetrieveFilters(call_tags=["x"])
        elif kind == "empty_tags":
            filters = RetrieveFilters(call_tags=[])
        elif kind == "filters_only_external":
            filters = RetrieveFilters(external_id="abc")
        assert retrieve._choose_dense_mode(case["rows"], filters, call_ids) == case["mode"], case
        assert retrieve._dense_has_scoping(filters, call_ids) == case["scoped"], case


# ---------------------------------------------------------------- small pure helpers vs goldens
def test_vector_literal_and_parse(pure):
    for case in pure["vector_literal"]:
        vals = [float.fromhex(h) for h in case["values_hex"]]
        assert retrieve._vector_literal(vals) == case["literal"]
        parsed = retrieve._query_vector(case["literal"])
        assert np.array_equal(parsed.view(np.uint32), np.array(vals, dtype=np.float32).view(np.uint32))
    with pytest.raises(pkg.DenseEngineError):
        retrieve._query_vector("1,2,3")


def test_build_debug_lane(pure):
    for case in pure["debug_lane"]:
        assert retrieve._build_debug_lane(case["rows"], case["id_field"]) == case["lane"]


def test_extract_tech_tokens_matches_reference_golden(pure):
    for case in pure["tech_tokens"]:
        assert lexical.extract_tech_tokens(case["text"]) == case["tokens"], case["text"]


def test_extract_tech_tokens_reference_unit_cases():
    # tests/unit/test_ingest_utils.py:12-28 style known answers
    toks = lexical.extract_tech_tokens("Which ticket tracked the ECONNRESET issue for ABC-123 on v1.2.3?")
    assert toks == ["ABC-123", "ECONNRESET", "v1.2.3"]


def test_tech_index_matches_port():
    rng = np.random.default_rng(3)
    n = 400
    vocab = [f"T{i}" for i in range(30)] + ["t1", "ABC-1"]
    row_tokens = [list(rng.choice(vocab, size=int(rng.integers(0, 4)), replace=False)) for _ in range(n)]
    ids = np.arange(1, n + 1, dtype=np.int64) * 3
    started = (rng.integers(0, 20, size=n) * 1000).astype(np.int64)
    slots = rng.integers(0, 10, size=n).astype(np.int32)
    tags = (np.uint64(1) << rng.integers(0, 8, size=n).astype(np.uint64))
    cols = {"ids": ids, "started_at": started, "call_slot": slots, "tag_bits": tags}
    idx = lexical.TechTokenIndex()
    for r, toks in enumerate(row_tokens):
        idx.add_row(r, toks)
    for tokens, spec in [(["T1", "T2"], {}), (["t1"], {}), (["T3"], {"call_slots": [1, 2, 3]}),
                         (["T4", "T5", "nope"], {"date_from": 5000, "date_to": 15000}),
                         (["T6"], {"tag_mask": 0b1010}), (["T7"], {"call_slots": []}), ([], {})]:
        keep = ports.filter_rows(slots, started, tags, None, call_slots=spec.get("call_slots"),
                                 date_from_us=spec.get("date_from"), date_to_us=spec.get("date_to"),
                                 tag_mask=spec.get("tag_mask"))
        want = ports.tech_lane(row_tokens, ids, started, keep, tokens, 50)
        got = ids[idx.query(tokens, cols, 50, **spec)].tolist() if tokens else []
        assert got == want


# ---------------------------------------------------------------- embedding client (reference tests mirrored)
class _FakeResponse:
    def __init__(self, status_code, body, text=""):
        self.status_code, self._body, self.text = status_code, body, text

    def json(self):
        return self._body


class _FakeClient:
    calls = []
    response = None

    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def post(self, url, json=None):
        _FakeClient.calls.append((url, json))
        r = _FakeClient.response
        return r(json) if callable(r) else r


@pytest.fixture
def fake_http(monkeypatch):
    monkeypatch.setattr(settings, "embeddings_base_url", "http://embed.local/")
    monkeypatch.setattr(settings, "embeddings_dim", 4)
    monkeypatch.setattr(embeddings.httpx, "Client", _FakeClient)
    embeddings.set_embedder(None)
    _FakeClient.calls = []
    return _FakeClient


def test_embed_texts_posts_to_embed_and_validates(fake_http):
    fake_http.response = _FakeResponse(200, {"embeddings": [[1, 2, 3, 4], [5, 6, 7, 8]], "model": "m"})
    res = embeddings.embed_texts([" a ", "b", "  "])
    assert fake_http.calls[0][0] == "http://embed.local/embed"
    assert fake_http.calls[0][1]["texts"] == ["a", "b"]
    assert res.vectors == [[1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0]] and res.model == "m"


def test_embed_texts_rejects_wrong_dim_and_errors(fake_http):
    fake_http.response = _FakeResponse(200, {"embeddings": [[1, 2, 3]]})
    with pytest.raises(embeddings.EmbeddingClientError, match="has dim 3; expected 4"):
        embeddings.embed_texts(["a"])
    fake_http.response = _FakeResponse(500, {}, text="boom")
    with pytest.raises(embeddings.EmbeddingClientError, match="returned 500: boom"):
        embeddings.embed_texts(["a"])
    fake_http.response = _FakeResponse(200, {"embeddings": [[1, 2, 3, 4], [1, 2, 3, 4]]})
    with pytest.raises(embeddings.EmbeddingClientError, match="count mismatch"):
        embeddings.embed_texts(["a"])
    fake_http.response = _FakeResponse(200, {"nope": 1})
    with pytest.raises(embeddings.EmbeddingClientError, match="missing 'embeddings'"):
        embeddings.embed_texts(["a"])
    with pytest.raises(embeddings.EmbeddingClientError, match="at least one non-empty"):
        embeddings.embed_texts(["  "])


def test_embed_texts_batched_splits(fake_http):
    fake_http.response = lambda payload: _FakeResponse(
        200, {"embeddings": [[0, 0, 0, float(len(t))] for t in payload["texts"]]})
    res = embeddings.embed_texts_batched(["a", "bb", "ccc", "dddd", "eeeee"], batch_size=2)
    assert [len(c[1]["texts"]) for c in fake_http.calls] == [2, 2, 1]
    assert [v[3] for v in res.vectors] == [1.0, 2.0, 3.0, 4.0, 5.0]
    with pytest.raises(embeddings.EmbeddingClientError):
        embeddings.embed_texts_batched(["a"], batch_size=-1)


def test_embeddings_disabled_without_base_url(monkeypatch):
    monkeypatch.setattr(settings, "embeddings_base_url", "  ")
    embeddings.set_embedder(None)
    assert embeddings.embeddings_enabled() is False
    with pytest.raises(embeddings.EmbeddingClientError, match="EMBEDDINGS_BASE_URL is not configured"):
        embeddings.embed_texts(["a"])


# ---------------------------------------------------------------- C ABI: loads, exports everything declared
def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "cadence_dense.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cdr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    declared = _declared_symbols()
    assert len(declared) >= 20
    L = ctypes.CDLL(_ffi.library_path())
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/cadence_dense.h but not exported"
    # and the ctypes stub binds exactly the declared set
    assert sorted(_ffi.SIGNATURES) == declared
    assert _ffi.abi_version() == 1


def _sass_by_kernel():
    """SASS of the shipped library split by kernel name (cuobjdump -sass)."""
    out = subprocess.run(["cuobjdump", "-sass", _ffi.library_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    kernels, name = {}, None
    for line in out.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name is not None:
            kernels[name].append(line)
    return {k: "\n".join(v) for k, v in kernels.items()}


def test_library_is_sm100a_with_bulk_copy_and_no_legacy_mma():
    """The shipped binary is sm_100a only, the tensor-core lane is tcgen05 (UTCHMMA issued from TMA-fed shared memory,
    accumulators read back from TMEM with LDTM) with no legacy warp-level MMA anywhere, and the exact scan streams its
    tiles with bulk async copies.  profiles/r02/sass_summary.txt records the same counts."""
    out = subprocess.run(["cuobjdump", "-lelf", _ffi.library_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    elfs = [ln for ln in out.stdout.splitlines() if "ELF file" in ln]
    assert elfs and all("sm_100a" in ln for ln in elfs), elfs
    sass = _sass_by_kernel()
    gemm = [v for k, v in sass.items() if "gemm_topk_kernel" in k]
    scan = [v for k, v in sass.items() if "exact_scan_kernel" in k]
    assert len(gemm) >= 3 and len(scan) >= 8
    for body in gemm:
        assert "UTCHMMA" in body, "tcgen05.mma missing from a gemm_topk_kernel instantiation"
        assert "LDTM" in body, "tcgen05.ld missing"
        assert "UTMALDG" in body, "TMA tensor loads missing"
        assert "FMNMX3" in body, "3-input max missing from the epilogue"
    for body in scan:
        assert "UBLKCP" in body, "bulk async copy missing from an exact_scan_kernel instantiation"
    # the latency finalize hands data between the CTAs of its cluster with st.async (STAS) + mbarrier waits, not cluster barriers
    fin = [v for k, v in sass.items() if "scan_finalize_cluster_kernel" in k]
    assert len(fin) >= 5
    for body in fin:
        assert "STAS" in body and "SYNCS" in body, "st.async / mbarrier hand-over missing from the latency finalize"
    legacy = re.compile(r"\b(HMMA|IMMA|DMMA|HGMMA|QGMMA)\b")
    for name, body in sass.items():
        assert not legacy.search(body), f"legacy MMA instruction in {name}"


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present: covered by the gpu tier")
    with pytest.raises(pkg.DenseEngineError) as exc:
        _ffi.require_device()
    assert exc.value.code == _ffi.CDR_ERR_NO_DEVICE
    from cadence_rag_b200.store import DenseStore
    with pytest.raises(pkg.DenseEngineError):
        DenseStore("chunks", 16)


def test_product_never_imports_oracle():
    for dirpath, _d, files in os.walk(os.path.join(ROOT, "cadence_rag_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "liboracle" not in src, fn


# ---------------------------------------------------------------- write side (reference: tests/unit/test_embedding_pipeline.py)
def test_infer_batch_size_limit_from_triton_error():
    from cadence_rag_b200 import embedding_pipeline
    message = "Triton infer failed: [400] inference request batch-size must be <= 8 for 'qwen3_embed_4b_onnx'"
    assert embedding_pipeline.infer_batch_size_limit(message) == 8
    assert embedding_pipeline.infer_batch_size_limit("Maximum batch size exceeded: 16") == 16
    assert embedding_pipeline.infer_batch_size_limit("upstream unavailable") is None
    assert embedding_pipeline.infer_batch_size_limit("") is None


def test_embed_texts_adaptive_reduces_batch_size(monkeypatch):
    from cadence_rag_b200 import embedding_pipeline
    from cadence_rag_b200.embeddings import EmbeddingClientError, EmbeddingResult
    calls = []

    def _mock_embed_texts(texts):
        calls.append(len(texts))
        if len(texts) > 2:
            raise EmbeddingClientError("inference request batch-size must be <= 2")
        return EmbeddingResult(vectors=[[0.1] * embedding_pipeline.settings.embeddings_dim for _ in texts], model="mock-embed")

    monkeypatch.setattr(embedding_pipeline, "embed_texts", _mock_embed_texts)
    result = embedding_pipeline._embed_texts_adaptive(["a", "b", "c", "d", "e"], batch_size=5)
    assert calls == [5, 2, 2, 1]
    assert len(result.vectors) == 5 and result.model == "mock-embed"


def test_embed_texts_adaptive_raises_when_single_row_fails(monkeypatch):
    from cadence_rag_b200 import embedding_pipeline
    from cadence_rag_b200.embeddings import EmbeddingClientError

    def _mock_embed_texts(texts):
        raise EmbeddingClientError("upstream unavailable")

    monkeypatch.setattr(embedding_pipeline, "embed_texts", _mock_embed_texts)
    with pytest.raises(EmbeddingClientError):
        embedding_pipeline._embed_texts_adaptive(["only-one"], batch_size=4)


def test_run_embedding_backfill_argument_checks(monkeypatch):
    from cadence_rag_b200 import embedding_pipeline
    monkeypatch.setattr(embedding_pipeline, "embeddings_enabled", lambda: False)
    with pytest.raises(RuntimeError, match="EMBEDDINGS_BASE_URL"):
        embedding_pipeline.run_embedding_backfill([], batch_size=8)
    monkeypatch.setattr(embedding_pipeline, "embeddings_enabled", lambda: True)
    with pytest.raises(RuntimeError, match="EMBEDDINGS_BATCH_SIZE must be > 0"):
        embedding_pipeline.run_embedding_backfill([], batch_size=0)
    assert embedding_pipeline.run_embedding_backfill([], batch_size=4).rows_updated == 0


def test_batch_lane_choice_inside_mode_ann():
    """A batch in mode "ann" runs on the tensor-core lane unless the filter is selective enough for the exact
    lane's gather launch to beat a full multiplication (cadence_rag_b200.retrieve._batch_lane_is_faster)."""
    from cadence_rag_b200.retrieve import _batch_lane_is_faster

    class _S:
        rows, dim = 10_000_000, 1024
    assert _batch_lane_is_faster(_S, 1024, 10_000_000)            # unfiltered: 17 ms vs 342 full scans
    assert _batch_lane_is_faster(_S, 1024, 500_000)               # 5 % of the rows: still the batch lane
    assert not _batch_lane_is_faster(_S, 1024, 2_000)             # the planner's 2 000-row scoped sets
    assert not _batch_lane_is_faster(_S, 16, 100_000)             # small batch, 1 % of the rows
    assert _batch_lane_is_faster(_S, 128, 10_000_000)             # HBM-bound small batch, unfiltered


def test_dense_lane_of_a_request_group(monkeypatch):
    """Which kernel serves a group's dense lane inside the fused call (retrieve._group_dense_lane): scoped groups
    always the exact scan (their mode depends on COUNT(*)); unscoped ones plan "ann" -- one request scans the bf16
    rows, fewer than cadence_gpu_ann_min_batch share passes over the bf16 rows, >= that many take the tensor-core lane when it is the
    faster one; stores without bf16 rows, other widths and the switches fall back to the exact scan."""
    from cadence_rag_b200 import _ffi
    from cadence_rag_b200.retrieve import RetrieveFilters, _group_dense_lane
    EXACT, BATCH, SCAN = _ffi.CDR_DENSE_LANE_EXACT_F32, _ffi.CDR_DENSE_LANE_BATCH_BF16, _ffi.CDR_DENSE_LANE_SCAN_BF16

    class _Big:
        rows, dim, has_bf16, has_fp32 = 1_000_000, 1024, True, True

    class _Small(_Big):
        rows = 10_000                                              # 41 MB of rows: the cost model keeps the exact scan

    class _NoBf16(_Big):
        has_bf16 = False

    class _Wide(_Big):
        dim = 1536

    monkeypatch.setattr(settings, "cadence_gpu_ann_min_batch", 4)
    monkeypatch.setattr(settings, "cadence_gpu_ann_bf16_scan", 1)
    assert [_group_dense_lane(_Big, None, None, n) for n in (1, 2, 3, 4, 64)] == [SCAN, SCAN, SCAN, BATCH, BATCH]
    assert _group_dense_lane(_Big, RetrieveFilters(), None, 1) == SCAN             # an empty filter object scopes nothing
    scoped = RetrieveFilters(call_ids=["c1"])
    assert [_group_dense_lane(_Big, scoped, ["c1"], n) for n in (1, 64)] == [EXACT, EXACT]
    assert _group_dense_lane(_Big, RetrieveFilters(call_tags=["vip"]), None, 64) == EXACT
    assert _group_dense_lane(_Big, RetrieveFilters(external_id="x"), [], 64) == EXACT   # resolved to no call: scoped to nothing
    assert [_group_dense_lane(_Small, None, None, n) for n in (1, 64)] == [SCAN, EXACT]
    assert [_group_dense_lane(_NoBf16, None, None, n) for n in (1, 64)] == [EXACT, EXACT]
    assert [_group_dense_lane(_Wide, None, None, n) for n in (1, 2)] == [EXACT, EXACT]
    monkeypatch.setattr(settings, "cadence_gpu_ann_bf16_scan", 0)
    assert _group_dense_lane(_Big, None, None, 1) == EXACT
    monkeypatch.setattr(settings, "cadence_gpu_ann_min_batch", 128)
    assert _group_dense_lane(_Big, None, None, 64) == EXACT


# ---------------------------------------------------------------- /retrieve response contract vs the reference's own code
def _ref_request_models():
    """pydantic request models with the reference's fields and defaults (app/schemas.py:71-93); when the reference
    tree is present (this container) the REAL classes are used instead."""
    try:
        sys.path.insert(0, "/root/reference")
        from app.schemas import Budget as B, RetrieveRequest as RR       # noqa: WPS433
        return B, RR
    except Exception:
        from typing import Literal, Optional
        from pydantic import BaseModel, Field

        class B(BaseModel):
            max_evidence_items: int = 8
            max_total_chars: int = 6000

        class RR(BaseModel):
            query: str
            intent: Literal["auto", "decision", "action_items", "who_said", "troubleshooting", "status"] = "auto"
            filters: Optional[dict] = None
            budget: B = Field(default_factory=B)
            return_style: Literal["evidence_pack_json", "ids_only"] = "evidence_pack_json"
            debug: bool = False
        return B, RR
    finally:
        if sys.path and sys.path[0] == "/root/reference":
            sys.path.pop(0)


_RefBudget, _RefRetrieveRequest = _ref_request_models()


def test_retrieve_evidence_payload_form_needs_an_engine():
    from cadence_rag_b200 import retrieve as R
    R.set_default_engine(None)
    with pytest.raises(pkg.DenseEngineError):
        R.retrieve_evidence(_RefRetrieveRequest(query="anything"))


def test_retrieve_evidence_matches_reference_golden(monkeypatch, golden_dir):
    """tests/golden/reference_evidence.json holds the REFERENCE's retrieve_evidence (app/retrieve.py:392-688, run
    live by tests/golden/make_golden_evidence.py) on canned lane rows.  The same lanes replayed through
    cadence_rag_b200.retrieve.retrieve_evidence must give the same response: fused order, ids_only combine,
    budgeted evidence pack (clipping, <= 2 artifacts, <= 2 quotes per call), planner label, notes, debug."""
    import contextlib
    from cadence_rag_b200 import retrieve as R
    from cadence_rag_b200.embeddings import EmbeddingClientError, EmbeddingResult
    from oracle import ports
    with open(os.path.join(golden_dir, "reference_evidence.json")) as f:
        gold = json.load(f)
    monkeypatch.setattr(settings, "embeddings_exact_scan_threshold", gold["settings"]["embeddings_exact_scan_threshold"])
    monkeypatch.setattr(settings, "embeddings_hnsw_ef_search", gold["settings"]["embeddings_hnsw_ef_search"])
    tables = {"chunks": {r["chunk_id"]: r for r in gold["corpus"]["chunks"]},
              "artifact_chunks": {r["artifact_chunk_id"]: r for r in gold["corpus"]["artifacts"]}}

    class _Store:
        def __init__(self, name, key):
            rows = tables[name]
            self.table_name, self.key_field = name, key
            ids = np.array(sorted(rows), dtype=np.int64)
            calls = sorted({rows[i]["call_id"] for i in rows})
            self.call_ids_by_slot = calls
            self._cols = {"ids": ids, "call_slot": np.array([calls.index(rows[int(i)]["call_id"]) for i in ids], dtype=np.int32)}
            self.payload = {i: {k: v for k, v in rows[i].items() if k not in (key, "call_id")} for i in rows}

        def host_columns(self):
            return self._cols

    class _Engine:
        stores = {"chunks": _Store("chunks", "chunk_id"), "artifact_chunks": _Store("artifact_chunks", "artifact_chunk_id")}

        @contextlib.contextmanager
        def connect(self):
            yield object()

    def rows_of(case, lane, table, key):
        out = []
        for ident, score in case["lanes"][lane]:
            row = dict(tables[table][ident])
            if score is not None:
                row["score"] = score
            out.append(row)
        return out

    monkeypatch.setattr(R, "_fused_path_ok", lambda *a, **k: False)
    monkeypatch.setattr(R, "_rrf_merge", lambda lanes, key, k=60: ports.rrf_merge(lanes, key, k))   # K5 needs a GPU; the port is pinned to the reference
    n_pack = n_ids = 0
    for item in gold["cases"]:
        case, want = item["case"], item["response"]
        monkeypatch.setattr(R, "embeddings_enabled", lambda c=case: c["dense_enabled"])

        def _embed(texts, c=case):
            if c["embed_error"]:
                raise EmbeddingClientError(c["embed_error"])
            return EmbeddingResult(vectors=[[0.5] * 4], model="golden-embedder")
        monkeypatch.setattr(R, "embed_texts", _embed)
        monkeypatch.setattr(R, "_resolve_call_ids", lambda conn, filters, c=case: ["c1"] if c["scoped"] else None)
        monkeypatch.setattr(R, "_fetch_chunks_tech", lambda *a, c=case: rows_of(c, "tech_chunks", "chunks", "chunk_id"))
        monkeypatch.setattr(R, "_fetch_artifacts_tech", lambda *a, c=case: rows_of(c, "tech_artifacts", "artifact_chunks", "artifact_chunk_id"))
        monkeypatch.setattr(R, "_fetch_chunks_dense", lambda *a, c=case: rows_of(c, "dense_chunks", "chunks", "chunk_id"))
        monkeypatch.setattr(R, "_fetch_artifacts_dense", lambda *a, c=case: rows_of(c, "dense_artifacts", "artifact_chunks", "artifact_chunk_id"))
        monkeypatch.setattr(R, "_estimate_dense_candidates", lambda conn, table, filters, cids, c=case: c["candidates"][table])
        got = R.retrieve_evidence(_Engine(), case["query"], None, R.Budget(*case["budget"]), intent=case["intent"],
                                  return_style=case["return_style"], debug=case["debug"],
                                  bm25_chunks=rows_of(case, "bm25_chunks", "chunks", "chunk_id"),
                                  bm25_artifacts=rows_of(case, "bm25_artifacts", "artifact_chunks", "artifact_chunk_id"))
        got = json.loads(json.dumps(got, default=str))
        # the reference's own call shape, retrieve_evidence(payload: RetrieveRequest) (app/retrieve.py:392), with
        # pydantic request objects shaped like app/schemas.py:71-93, must give the same response
        R.set_default_engine(eng := _Engine())
        eng.bm25_lanes = lambda q, f, c=case: (rows_of(c, "bm25_chunks", "chunks", "chunk_id"),
                                               rows_of(c, "bm25_artifacts", "artifact_chunks", "artifact_chunk_id"))
        payload = _RefRetrieveRequest(query=case["query"], intent=case["intent"], return_style=case["return_style"],
                                      debug=case["debug"], budget=_RefBudget(max_evidence_items=case["budget"][0],
                                                                             max_total_chars=case["budget"][1]))
        got2 = json.loads(json.dumps(R.retrieve_evidence(payload), default=str))
        R.set_default_engine(None)
        got2.pop("query_id"); g1 = dict(got); g1.pop("query_id")
        assert got2 == g1, case["query"]
        assert set(want) - {"debug"} <= set(got), (case["query"], set(want) - set(got))
        for key in want:
            if key == "query_id":
                continue
            if key == "debug":
                for part in ("lanes", "limits", "dense"):
                    assert got["debug"][part] == want["debug"][part], (case["return_style"], part)
                continue
            assert got[key] == want[key], (case["return_style"], case["budget"], key)
        if case["return_style"] == "ids_only":
            n_ids += 1
        else:
            n_pack += 1
    assert n_pack >= 20 and n_ids >= 5


def test_resolve_call_ids_and_filter_clause_match_reference_golden(golden_dir):
    """_resolve_call_ids (app/retrieve.py:46-90) and _build_filter_clause (:93-120), run live by
    tests/golden/make_golden_evidence.py against a stand-in `calls` table, vs our external-id map / _filter_spec."""
    from uuid import UUID
    from cadence_rag_b200.retrieve import DenseEngine, _filter_spec, _resolve_call_ids
    with open(os.path.join(golden_dir, "reference_evidence.json")) as f:
        gold = json.load(f)
    eng = DenseEngine()
    for call_id, ext, src in gold["resolve_call_ids"]["calls"]:
        eng.register_call(UUID(call_id), external_id=ext, external_source=src)
    for case in gold["resolve_call_ids"]["cases"]:
        spec = case["filters"]
        filters = None
        if spec is not None:
            spec = dict(spec)
            if "call_ids" in spec:
                spec["call_ids"] = [UUID(c) for c in spec["call_ids"]]
            filters = RetrieveFilters(**spec)
        with eng.connect() as conn:
            got = _resolve_call_ids(conn, filters)
        assert (None if got is None else [str(c) for c in got]) == case["call_ids"], case

    class _Store:
        synthetic = None
        call_ids_by_slot = ["x", "y"]

        def slot_of_call(self, c, create=False):
            return {"x": 0, "y": 1}.get(c)

        def bits_of_tags(self, tags, create=False):
            return sum(1 << {"a": 0, "b": 1}.get(t, 63) for t in tags if t in ("a", "b"))

        def tag_filter(self, tags):
            return self.bits_of_tags(tags), None

    for case in gold["filter_clause"]:
        spec = case["filters"]
        filters = None
        if spec is not None:
            spec = {k: (datetime.fromisoformat(v) if k.startswith("date_") else v) for k, v in spec.items()}
            filters = RetrieveFilters(**spec)
        ours = _filter_spec(_Store(), filters, case["call_ids"])
        keys = set(case["param_keys"])
        assert (ours["date_from"] is not None) == ("date_from" in keys), case
        assert (ours["date_to"] is not None) == ("date_to" in keys), case
        assert (ours["call_slots"] is not None) == ("call_ids" in keys), case
        assert (ours["tag_mask"] is not None) == ("call_tags" in keys) == case["join_calls"], case
        assert (case["where"] == "TRUE") == all(v is None for v in ours.values())


def test_embedding_f32_equals_the_literal_round_trip():
    """retrieve._embedding_f32 == CAST(_vector_literal(v) AS vector): the float32-representable shortcut and the
    text path agree, for float32 inputs, fp64 decimals (a JSON body), tiny / huge magnitudes and NaN."""
    rng = np.random.default_rng(9)
    f32 = rng.standard_normal(1024).astype(np.float32)
    f64 = rng.standard_normal(1024) * 10.0 ** rng.integers(-12, 6, size=1024)
    for values in (f32.astype(np.float64).tolist(), f64.tolist(), [0.1, 1 / 3, 1e-12, -0.0, 1.0, 3.4e38, 1e-45, float("nan")]):
        via_text = retrieve._query_vector(retrieve._vector_literal(values))
        got = retrieve._embedding_f32(values)
        assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), via_text.view(np.uint32)) or \
            (np.isnan(got).any() and np.array_equal(np.isnan(got), np.isnan(via_text)))
    assert np.array_equal(retrieve._embedding_f32(f32.astype(np.float64).tolist()), f32)
    # batch form: row by row the same bits, mixed exact / inexact / NaN rows
    rows = [f32.astype(np.float64).tolist(), f64.tolist(), [float("nan")] + f64.tolist()[1:], (f32 * 3).astype(np.float64).tolist()]
    many = retrieve._embeddings_f32(rows)
    assert many.dtype == np.float32 and many.shape == (4, 1024)
    for got, values in zip(many, rows):
        assert np.array_equal(got.view(np.uint32), retrieve._embedding_f32(values).view(np.uint32))


def test_request_batcher_queueing(monkeypatch):
    """RequestBatcher host logic without a GPU: concurrent clients are served in batches by one retrieve_ids_batch
    call each, every client gets its own response (debug payload only if it asked), a failing request gets its own
    error (the batch is retried member by member), and a closed batcher refuses work."""
    import threading
    import time
    from cadence_rag_b200 import retrieve as R
    seen = []

    def fake_batch(engine, queries, filters, bm25_chunks=None, bm25_artifacts=None, debug=False):
        seen.append((list(queries), list(filters), debug))
        time.sleep(0.01)                                   # the "GPU": lets the queue fill behind the worker
        if any(q == "boom" for q in queries):
            raise R.DenseEngineError("engine failed")
        return [{"retrieved_ids": [f"chunk:{q}:{f}"], **({"debug": {"q": q}} if debug else {})} for q, f in zip(queries, filters)]

    def fake_one(engine, query, filters=None, bm25_chunks=(), bm25_artifacts=(), debug=False):
        return fake_batch(engine, [query], [filters], debug=debug)[0]     # the per-member retry after a failed batch

    monkeypatch.setattr(R, "retrieve_ids_batch", fake_batch)
    monkeypatch.setattr(R, "retrieve_ids", fake_one)
    engine = type("E", (), {"stores": {}})()
    batcher = R.RequestBatcher(engine, max_batch=8, max_wait_s=5e-3)
    results, errors = {}, {}

    def client(t):
        for j in range(5):
            try:
                results[(t, j)] = batcher.retrieve_ids(f"q{t}-{j}", t, debug=(t % 2 == 0))
            except R.DenseEngineError as exc:
                errors[(t, j)] = str(exc)
    threads = [threading.Thread(target=client, args=(t,)) for t in range(12)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert len(results) == 60 and not errors
    for (t, j), resp in results.items():
        assert resp["retrieved_ids"] == [f"chunk:q{t}-{j}:{t}"]
        assert ("debug" in resp) == (t % 2 == 0)
    assert batcher.requests_served == 60 and batcher.batches_served < 40 and max(len(b[0]) for b in seen) <= 8
    # a failing batch is retried member by member: the request that fails sees its error, the batcher keeps serving
    with pytest.raises(R.DenseEngineError, match="engine failed"):
        batcher.retrieve_ids("boom", None)
    assert batcher.retrieve_ids("after", 7)["retrieved_ids"] == ["chunk:after:7"]
    batcher.close()
    with pytest.raises(R.DenseEngineError):
        batcher.retrieve_ids("late", None)


# ---------------------------------------------------------------- offline evaluation (eval/run_eval.py) on the engine's output
def test_eval_metrics_match_reference_golden(golden_dir):
    """recall@k / MRR / nDCG@k of `eval_replay.compute_metrics` equal the reference's `compute_metrics`
    (eval/run_eval.py:26-65, run live by tests/golden/make_golden_eval.py) bit for bit, key order included."""
    from cadence_rag_b200.eval_replay import compute_metrics
    with open(os.path.join(golden_dir, "reference_eval_metrics.json")) as f:
        gold = json.load(f)
    assert len(gold["cases"]) >= 20
    for case in gold["cases"]:
        got = compute_metrics(case["gold"], case["results"], case["ks"])
        assert list(got.keys()) == case["metric_order"]
        assert {k: float(v).hex() for k, v in got.items()} == case["metrics_hex"]


def test_eval_cli_and_jsonl_round_trip(tmp_path, capsys):
    from cadence_rag_b200 import eval_replay as E
    gold_rows = [{"query_id": "a", "query": "x", "relevant_ids": ["chunk:1", "chunk:2"]},
                 {"query_id": "b", "query": "y", "relevant_ids": []},
                 {"query_id": "c", "query": "z", "relevant_ids": ["artifact_chunk:9"]}]
    result_rows = [{"query_id": "a", "retrieved_ids": ["chunk:7", "chunk:2", "chunk:1"]},
                   {"query_id": "c", "retrieved": ["artifact_chunk:9"]}]           # legacy key, run_eval.py:82
    gold_path, res_path = str(tmp_path / "gold.jsonl"), str(tmp_path / "res.jsonl")
    E.dump_jsonl(gold_rows, gold_path)
    E.dump_jsonl(result_rows, res_path)
    assert E.load_jsonl(gold_path) == gold_rows
    E.main(["--gold", gold_path, "--results", res_path, "--k", "1", "3"])
    m = json.loads(capsys.readouterr().out)
    assert m["recall@1"] == 0.5 and m["recall@3"] == 1.0          # (0 + 1)/2, (1 + 1)/2; query b is skipped
    assert m["mrr"] == 0.75                                        # (1/2 + 1)/2
    assert E.compute_metrics({}, {}, [5]) == {"recall@5": 0.0, "mrr": 0.0, "ndcg@5": 0.0}


# ---------------------------------------------------------------- the fused request path above the C call (no GPU)
class _FakeFusedStore:
    """Stands where DenseStore stands for the fused path: `hybrid_retrieve` answers from a deterministic toy model
    (dense score of row r for a query = -(|sum(q) * 1000 - r| mod 997), filters keep every `stride`-th row) and fuses
    the lanes with the restated RRF, in the array layout of the real call.  It records the specs it was given."""
    key_field, dim, has_fp32, has_bf16, rows, synthetic, device = "chunk_id", 256, True, True, 5_000_000, None, 0
    table_name = "chunks"

    def __init__(self):
        self.calls = []
        self.fail_dense = False
        self.fail_code = _ffi.CDR_ERR_NO_DEVICE

    def slot_of_call(self, call_id, create=False):
        return int(call_id)

    def bits_of_tags(self, tags, create=False):
        return 1

    def _dense(self, q, spec, k):
        stride = 1 if not spec or spec.get("call_slots") is None else 1 + len(spec["call_slots"])
        base = int(abs(float(np.sum(q))) * 1000) % 4000
        rows = [r for r in range(base, base + 400) if r % stride == 0]
        rows.sort(key=lambda r: ((r * 7919) % 997, r))
        return [(r + 1, 1.0 - ((r * 7919) % 997) / 1000.0) for r in rows[:k]]

    def hybrid_retrieve(self, queries, dense_k, *, tech_index=None, token_ids=None, n_tokens=None, tech_limit=50,
                        bm25_ids=None, bm25_offsets=None, rrf_k=60, filter_spec=None, max_out=None, filter_specs=None,
                        group_offsets=None):
        self.calls.append({"filter_spec": filter_spec, "filter_specs": filter_specs, "group_offsets": group_offsets,
                           "nq": 0 if queries is None else len(queries)})
        if self.fail_dense and queries is not None:
            raise _ffi.DenseEngineError("device lost", self.fail_code)
        nq = len(bm25_offsets) - 1
        specs = [filter_spec] * nq
        counts = 1234
        if filter_specs is not None:
            specs = [filter_specs[g] for g in range(len(filter_specs)) for _ in range(group_offsets[g], group_offsets[g + 1])]
            counts = [1000 + g for g in range(len(filter_specs))]
        kd = dense_k if queries is not None else 0
        width = max(1, int(np.diff(bm25_offsets).max()) + tech_limit + kd)
        out = {"dense_ids": np.full((nq, max(kd, 1)), -1, dtype=np.int64), "dense_scores": np.zeros((nq, max(kd, 1))),
               "dense_n": np.zeros(nq, dtype=np.int32), "tech_ids": np.full((nq, tech_limit), -1, dtype=np.int64),
               "tech_n": np.zeros(nq, dtype=np.int32), "fused_ids": np.full((nq, width), -1, dtype=np.int64),
               "fused_scores": np.zeros((nq, width)), "fused_mask": np.zeros((nq, width), dtype=np.uint32),
               "fused_n": np.zeros(nq, dtype=np.int32), "count": counts}
        for i in range(nq):
            lanes = {"bm25": [{"chunk_id": int(v)} for v in bm25_ids[bm25_offsets[i]:bm25_offsets[i + 1]]], "tech_tokens": []}
            if queries is not None:
                hits = self._dense(queries[i], specs[i], dense_k)
                out["dense_n"][i] = len(hits)
                for j, (cid, sc) in enumerate(hits):
                    out["dense_ids"][i, j], out["dense_scores"][i, j] = cid, sc
                lanes["dense"] = [{"chunk_id": cid} for cid, _ in hits]
            fused = ports.rrf_merge(lanes, "chunk_id", rrf_k)
            out["fused_n"][i] = len(fused)
            for j, (row, hit, sc) in enumerate(fused):
                out["fused_ids"][i, j], out["fused_scores"][i, j] = row["chunk_id"], sc
                out["fused_mask"][i, j] = sum(1 << l for l, name in enumerate(("bm25", "tech_tokens", "dense")) if name in hit)
        return out


def test_fused_request_path_above_the_c_call(monkeypatch):
    """Everything `retrieve_ids` / `retrieve_ids_batch` do around the fused C call, against a toy store: the dict-free
    response (no debug) equals the row-dict response, per-request filters are grouped and the responses return in
    request order, unscoped groups get their dense lane, the dense lane fails open on both kinds of error."""
    monkeypatch.setattr(settings, "embeddings_dim", 256)
    monkeypatch.setattr(settings, "embeddings_base_url", "http://embedder")
    monkeypatch.setattr(settings, "cadence_gpu_ann_min_batch", 4)
    monkeypatch.setattr(settings, "cadence_gpu_ann_bf16_scan", 1)
    store = _FakeFusedStore()
    eng = retrieve.DenseEngine()
    eng.stores["chunks"] = store
    vec = lambda t: [float((sum(map(ord, t)) * (j + 3)) % 17) / 16.0 for j in range(256)]     # noqa: E731
    embeddings.set_embedder(lambda batch: embeddings.EmbeddingResult(vectors=[vec(t) for t in batch], model="toy"))
    bm25 = [{"chunk_id": 11, "text": "a"}, {"chunk_id": 3012, "text": "b"}]
    try:
        scoped = RetrieveFilters(call_ids=[1, 2])
        texts = [f"question {i}" for i in range(9)]
        filt = [None, scoped, None, None, scoped, None, RetrieveFilters(call_ids=[7]), None, None]
        full = [retrieve.retrieve_ids(eng, t, f, bm25_chunks=bm25 if i % 2 else (), debug=True) for i, (t, f) in enumerate(zip(texts, filt))]
        assert all(c["nq"] == 1 for c in store.calls)
        lanes = [(c["filter_spec"] or {}).get("dense_lane", 0) for c in store.calls]
        assert lanes == [(_ffi.CDR_DENSE_LANE_SCAN_BF16 if f is None else 0) for f in filt]       # single unscoped: bf16 scan
        lean = [retrieve.retrieve_ids(eng, t, f, bm25_chunks=bm25 if i % 2 else ()) for i, (t, f) in enumerate(zip(texts, filt))]
        assert lean == [{"retrieved_ids": r["retrieved_ids"]} for r in full]
        assert full[1]["debug"]["dense"]["modes"]["chunks"] == "exact" and full[0]["debug"]["dense"]["modes"]["chunks"] == "ann"
        assert full[1]["debug"]["dense"]["candidate_rows"]["chunks"] == 1234
        assert full[1]["debug"]["lanes"]["chunks"]["bm25"][0] == {"chunk_id": 11, "rank": 1, "score": None}
        # one fused call for all nine: three groups (unscoped x6 -> tensor-core lane, the two scoped filters -> exact)
        store.calls.clear()
        many = retrieve.retrieve_ids_batch(eng, texts, filt, bm25_chunks=[bm25 if i % 2 else [] for i in range(9)], debug=True)
        assert len(store.calls) == 1 and store.calls[0]["nq"] == 9
        call = store.calls[0]
        assert call["group_offsets"] == [0, 6, 8, 9]
        assert [(s or {}).get("dense_lane", 0) for s in call["filter_specs"]] == [_ffi.CDR_DENSE_LANE_BATCH_BF16, 0, 0]
        assert [(s or {}).get("call_slots") for s in call["filter_specs"]] == [None, [1, 2], [7]]
        for i in range(9):
            assert many[i]["retrieved_ids"] == full[i]["retrieved_ids"], i
            assert many[i]["debug"]["lanes"] == full[i]["debug"]["lanes"] and many[i]["debug"]["fused"] == full[i]["debug"]["fused"]
        assert [m["debug"]["dense"]["candidate_rows"]["chunks"] for m in many] == [1000, 1001, 1000, 1000, 1001, 1000, 1002, 1000, 1000]
        assert retrieve.retrieve_ids_batch(eng, texts, filt, bm25_chunks=[bm25 if i % 2 else [] for i in range(9)]) == lean
        # blank queries keep their slot
        mixed = retrieve.retrieve_ids_batch(eng, ["  ", texts[0], ""], None)
        assert mixed[0] == {"retrieved_ids": []} and mixed[2] == {"retrieved_ids": []} and mixed[1] == lean[0]
        # the embedding service fails: lexical-only, dense_error carried in the debug payload
        def boom(batch):
            raise embeddings.EmbeddingClientError("embedding HTTP request failed: down")
        embeddings.set_embedder(boom)
        off = retrieve.retrieve_ids(eng, texts[1], None, bm25_chunks=bm25, debug=True)
        assert off["retrieved_ids"] == ["chunk:11", "chunk:3012"] and off["debug"]["dense"]["enabled"] is False
        assert "down" in off["debug"]["dense"]["error"]
        assert retrieve.retrieve_ids(eng, texts[1], None, bm25_chunks=bm25) == {"retrieved_ids": ["chunk:11", "chunk:3012"]}
        # the engine fails on the dense lane: same fail-open, for the one-request and the batch form
        embeddings.set_embedder(lambda batch: embeddings.EmbeddingResult(vectors=[vec(t) for t in batch], model="toy"))
        store.fail_dense = True
        off = retrieve.retrieve_ids(eng, texts[1], None, bm25_chunks=bm25, debug=True)
        assert off["retrieved_ids"] == ["chunk:11", "chunk:3012"] and off["debug"]["dense"]["error"] == "device lost"
        assert retrieve.retrieve_ids_batch(eng, texts[:2], None, bm25_chunks=[bm25, []]) == [{"retrieved_ids": ["chunk:11", "chunk:3012"]}, {"retrieved_ids": []}]
        # ... but only for failures the lane can degrade from: a CUDA error / bad argument / bad state is an engine
        # fault and surfaces (the reference fails open on EmbeddingClientError alone)
        for code in (_ffi.CDR_ERR_CUDA, _ffi.CDR_ERR_INVALID, _ffi.CDR_ERR_STATE):
            store.fail_code = code
            with pytest.raises(_ffi.DenseEngineError):
                retrieve.retrieve_ids(eng, texts[1], None, bm25_chunks=bm25, debug=True)
            with pytest.raises(_ffi.DenseEngineError):
                retrieve.retrieve_ids_batch(eng, texts[:2], None, bm25_chunks=[bm25, []])
    finally:
        embeddings.set_embedder(None)


def test_request_batcher_over_the_fused_path(monkeypatch):
    """RequestBatcher with the real retrieve_ids_batch underneath (toy store, no GPU): 8 client threads with mixed
    filters get exactly the one-request responses, requests were served in shared fused calls, and every call's
    groups are consistent (offsets cover the batch, one spec per group)."""
    import threading
    monkeypatch.setattr(settings, "embeddings_dim", 256)
    monkeypatch.setattr(settings, "embeddings_base_url", "http://embedder")
    store = _FakeFusedStore()
    eng = retrieve.DenseEngine()
    eng.stores["chunks"] = store
    vec = lambda t: [float((sum(map(ord, t)) * (j + 3)) % 17) / 16.0 for j in range(256)]     # noqa: E731
    embeddings.set_embedder(lambda batch: embeddings.EmbeddingResult(vectors=[vec(t) for t in batch], model="toy"))
    filters = [None, RetrieveFilters(call_ids=[1, 2]), RetrieveFilters(call_ids=[7]), None]
    texts = [f"client question {i}" for i in range(12)]
    try:
        want = {(i, j): retrieve.retrieve_ids(eng, texts[i], filters[j], debug=True)["retrieved_ids"]
                for i in range(12) for j in range(4)}
        store.calls.clear()
        batcher = retrieve.RequestBatcher(eng, max_batch=16, max_wait_s=5e-3)
        got, errs = {}, []

        def client(t):
            try:
                for r in range(6):
                    i, j = (t * 5 + r) % 12, (t + r) % 4
                    got[(t, r)] = ((i, j), batcher.retrieve_ids(texts[i], filters[j])["retrieved_ids"])
            except Exception as exc:   # noqa: BLE001
                errs.append(repr(exc))
        threads = [threading.Thread(target=client, args=(t,)) for t in range(8)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        batcher.close()
        assert not errs, errs
        assert len(got) == 48 and all(ids == want[key] for key, ids in got.values())
        assert batcher.requests_served == 48 and batcher.batches_served < 48
        assert sum(c["nq"] for c in store.calls) == 48
        for c in store.calls:
            assert c["filter_specs"] is not None and c["group_offsets"][0] == 0 and c["group_offsets"][-1] == c["nq"]
            assert len(c["filter_specs"]) == len(c["group_offsets"]) - 1 <= 3
    finally:
        embeddings.set_embedder(None)


def test_request_batcher_isolates_a_failing_request(monkeypatch):
    """One client's malformed request must not fail the other clients of its batch: when the fused batch call raises,
    the batcher retries the members one at a time and each ticket gets its own response or its own error."""
    import threading
    monkeypatch.setattr(settings, "embeddings_dim", 256)
    monkeypatch.setattr(settings, "embeddings_base_url", "http://embedder")
    store = _FakeFusedStore()
    eng = retrieve.DenseEngine()
    eng.stores["chunks"] = store
    vec = lambda t: [float((sum(map(ord, t)) * (j + 3)) % 17) / 16.0 for j in range(256)]     # noqa: E731
    embeddings.set_embedder(lambda batch: embeddings.EmbeddingResult(vectors=[vec(t) for t in batch], model="toy"))
    bad_rows = [{"not_the_id_field": 1}]                                      # a BM25 row without chunk_id: KeyError
    try:
        want = retrieve.retrieve_ids(eng, "good question", None)
        batcher = retrieve.RequestBatcher(eng, max_batch=8, max_wait_s=0.05)
        results, errors = {}, {}

        def client(i):
            try:
                results[i] = batcher.retrieve_ids("good question", None, bm25_chunks=bad_rows if i == 3 else ())
            except Exception as exc:   # noqa: BLE001
                errors[i] = exc
        threads = [threading.Thread(target=client, args=(i,)) for i in range(6)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(30)
        batcher.close()
        assert sorted(errors) == [3] and isinstance(errors[3], KeyError)
        assert sorted(results) == [0, 1, 2, 4, 5] and all(r == want for r in results.values())
    finally:
        embeddings.set_embedder(None)
