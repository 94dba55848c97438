#!/usr/bin/env python
"""Generates tests/golden/reference_evidence.json: the REFERENCE's own `retrieve_evidence`
(app/retrieve.py:392-688, imported live through oracle/ref_stub.py) run on canned lane rows.

The reference reaches its lanes through SQL; here every lane function of app.retrieve
(_fetch_*_bm25 / _fetch_*_tech / _estimate_dense_candidates / _fetch_*_dense / _resolve_call_ids),
the embedding client and the connection are replaced by stand-ins that return the case's rows, so
what is exercised -- and pinned -- is everything AFTER the lanes: RRF, the ids_only combine, the
budgeted evidence pack (artifacts first, <= 2 quotes per call, snippet clipping, char budget), the
planner label and the notes / debug payloads.  tests/test_host_logic.py replays the same cases
through cadence_rag_b200.retrieve.retrieve_evidence.

    python tests/golden/make_golden_evidence.py        # build container only (needs /root/reference)
"""
import json
import os
import random
import sys
from contextlib import contextmanager

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_stub  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make_corpus(rng):
    """Payload rows of both tables (what the SQL SELECT lists return)."""
    chunks, artifacts = {}, {}
    for cid in range(1, 201):
        call = f"00000000-0000-0000-0000-{cid % 23:012d}"
        words = rng.choice([3, 8, 20, 40, 200])
        text = " ".join(rng.choice(["latency", "rollout", "ECONNRESET", "tiering", "budget", "SSD", "vs", "the", "a"])
                        for _ in range(words)) + ("   " if cid % 7 == 0 else "")
        chunks[cid] = {"chunk_id": cid, "call_id": call, "speaker": rng.choice(["alice", "bob", None]),
                       "start_ts_ms": cid * 1000, "end_ts_ms": cid * 1000 + 900, "text": text}
    for aid in range(1, 61):
        call = f"00000000-0000-0000-0000-{aid % 23:012d}"
        words = rng.choice([5, 30, 60, 220])
        artifacts[aid] = {"artifact_chunk_id": aid, "call_id": call, "artifact_id": 1000 + aid // 3,
                          "kind": rng.choice(["summary", "decisions", "action_items"]),
                          "content": " ".join(rng.choice(["decision", "owner", "deadline", "ABC-123", "v1.2.3", "to", "of"])
                                              for _ in range(words))}
    return chunks, artifacts


def lane(rng, table, universe, n, with_score):
    ids = rng.sample(sorted(universe), min(n, len(universe)))
    rows = []
    for rank, i in enumerate(ids):
        row = dict(table[i])
        if with_score:
            row["score"] = round(1.0 - rank * 0.01 - rng.random() * 0.001, 6)
        rows.append(row)
    return rows


def cases(rng, chunks, artifacts):
    out = []
    shapes = [  # (bm25_c, tech_c, dense_c, bm25_a, tech_a, dense_a)
        (50, 50, 50, 10, 50, 10), (0, 0, 50, 0, 0, 10), (50, 0, 0, 10, 0, 0), (3, 2, 5, 1, 0, 2), (0, 0, 0, 0, 0, 0),
        (20, 50, 50, 0, 3, 10), (50, 50, 50, 10, 10, 10), (1, 1, 1, 1, 1, 1),
    ]
    budgets = [(8, 6000), (3, 6000), (8, 900), (1, 100), (20, 100000), (8, 0), (0, 6000), (5, 1601)]
    for si, shape in enumerate(shapes):
        for bi, budget in enumerate(budgets):
            if (si + bi) % 3 and si > 1:
                continue
            dense_enabled = shape[2] + shape[5] > 0 or (si + bi) % 2 == 0
            overlap = rng.choice([40, 120, 200])
            cu = set(rng.sample(range(1, 201), overlap))
            au = set(rng.sample(range(1, 61), min(overlap, 60)))
            c = {"query": rng.choice(["Which ticket tracked the ECONNRESET issue for ABC-123 on v1.2.3?", "status of the rollout",
                                      "  object storage tiering vs SSD  "]),
                 "intent": rng.choice(["auto", "decision", "status"]), "budget": list(budget),
                 "return_style": "ids_only" if (si * 7 + bi) % 5 == 0 else "evidence_pack_json",
                 "debug": (si + bi) % 2 == 1, "dense_enabled": dense_enabled,
                 "embed_error": "embedding service returned 503: busy" if (si == 2 and bi == 1) else None,
                 "candidates": {"chunks": rng.choice([0, 1500, 2000, 2001, 90000]), "artifact_chunks": rng.choice([0, 10, 5000])},
                 "scoped": rng.choice([True, False]),
                 "lanes": {"bm25_chunks": lane(rng, chunks, cu, shape[0], True), "tech_chunks": lane(rng, chunks, cu, shape[1], False),
                           "dense_chunks": lane(rng, chunks, cu, shape[2], True), "bm25_artifacts": lane(rng, artifacts, au, shape[3], True),
                           "tech_artifacts": lane(rng, artifacts, au, shape[4], False),
                           "dense_artifacts": lane(rng, artifacts, au, shape[5], True)}}
            out.append(c)
    out.append({"query": "   ", "intent": "auto", "budget": [8, 6000], "return_style": "evidence_pack_json", "debug": False,
                "dense_enabled": True, "embed_error": None, "candidates": {"chunks": 0, "artifact_chunks": 0}, "scoped": False,
                "lanes": {k: [] for k in ("bm25_chunks", "tech_chunks", "dense_chunks", "bm25_artifacts", "tech_artifacts", "dense_artifacts")}})
    out.append(dict(out[-1], return_style="ids_only"))
    return out


def compact(case):
    c = dict(case)
    c["lanes"] = {name: [[r.get("chunk_id", r.get("artifact_chunk_id")), r.get("score")] for r in rows]
                  for name, rows in case["lanes"].items()}
    return c


def run_reference(ref, case):
    R = ref.retrieve
    L = case["lanes"]
    saved = {}

    def patch(name, value):
        saved[name] = getattr(R, name)
        setattr(R, name, value)

    class _Conn:
        pass

    class _Engine:
        @contextmanager
        def connect(self):
            yield _Conn()

    call_ids = ["c1"] if case["scoped"] else None
    patch("engine", _Engine())
    patch("uuid4", lambda: "00000000-0000-4000-8000-000000000000")
    patch("_resolve_call_ids", lambda conn, filters: call_ids)
    patch("_fetch_chunks_bm25", lambda *a: [dict(r) for r in L["bm25_chunks"]])
    patch("_fetch_artifacts_bm25", lambda *a: [dict(r) for r in L["bm25_artifacts"]])
    patch("_fetch_chunks_tech", lambda *a: [dict(r) for r in L["tech_chunks"]])
    patch("_fetch_artifacts_tech", lambda *a: [dict(r) for r in L["tech_artifacts"]])
    patch("_fetch_chunks_dense", lambda *a: [dict(r) for r in L["dense_chunks"]])
    patch("_fetch_artifacts_dense", lambda *a: [dict(r) for r in L["dense_artifacts"]])
    patch("_estimate_dense_candidates", lambda conn, table, filters, cids: case["candidates"][table])
    patch("embeddings_enabled", lambda: case["dense_enabled"])

    def _embed(texts):
        if case["embed_error"]:
            raise R.EmbeddingClientError(case["embed_error"])
        return ref.embeddings.EmbeddingResult(vectors=[[0.5] * 4], model="golden-embedder")
    patch("embed_texts", _embed)
    try:
        S = ref.schemas
        req = S.RetrieveRequest(query=case["query"], intent=case["intent"],
                                budget=S.Budget(max_evidence_items=case["budget"][0], max_total_chars=case["budget"][1]),
                                return_style=case["return_style"], debug=case["debug"])
        return R.retrieve_evidence(req)
    finally:
        for name, value in saved.items():
            setattr(R, name, value)


def resolve_cases(ref):
    """_resolve_call_ids (app/retrieve.py:46-90) against a stand-in `calls` table: the fake connection answers
    the function's two SELECTs from the bound parameters (external_source IS NOT DISTINCT FROM :external_source)."""
    from uuid import UUID
    calls = [(UUID(int=i + 1), f"ext-{i // 3}", [None, "crm", "zoom"][i % 3]) for i in range(18)]
    calls.append((UUID(int=100), "ext-2", "crm"))        # same (external_id, source) on two calls

    class _Result:
        def __init__(self, rows):
            self._rows = rows

        def fetchall(self):
            return self._rows

    class _Conn:
        def execute(self, _sql, params):
            rows = [(c,) for c, ext, src in calls
                    if ext == params["external_id"] and ("external_source" not in params or src == params["external_source"])]
            return _Result(rows)

    F = ref.schemas.RetrieveFilters
    specs = [None, {}, {"call_ids": []}, {"call_ids": [str(calls[4][0]), str(calls[0][0])]}, {"external_id": "ext-2"},
             {"external_id": "ext-2", "external_source": "crm"}, {"external_id": "ext-2", "external_source": "teams"},
             {"external_id": "missing"}, {"external_id": "ext-1", "call_ids": [str(calls[3][0]), str(calls[9][0])]},
             {"external_id": "ext-1", "call_ids": [str(calls[9][0])]}, {"external_id": "", "call_ids": [str(calls[2][0])]},
             {"external_source": "crm"}, {"call_tags": ["a"]}]
    out = []
    for spec in specs:
        filters = None if spec is None else F(**spec)
        got = ref.retrieve._resolve_call_ids(_Conn(), filters)
        out.append({"filters": spec, "call_ids": None if got is None else [str(c) for c in got]})
    return {"calls": [[str(c), ext, src] for c, ext, src in calls], "cases": out}


def filter_clause_cases(ref):
    """_build_filter_clause (app/retrieve.py:93-120): which predicates a filter produces."""
    from datetime import datetime, timezone
    F = ref.schemas.RetrieveFilters
    t = datetime(2026, 2, 9, tzinfo=timezone.utc)
    out = []
    for spec, call_ids in [(None, None), ({}, None), ({"date_from": t}, None), ({"date_to": t}, ["x"]), ({"call_tags": ["a", "b"]}, None),
                           ({"date_from": t, "date_to": t, "call_tags": ["z"]}, []), ({}, ["x", "y"]), ({"call_tags": []}, None),
                           (None, ["x"])]:
        filters = None if spec is None else F(**spec)
        where, params, join = ref.build_filter_clause(filters, "chunks", call_ids)
        out.append({"filters": None if spec is None else {k: (v.isoformat() if hasattr(v, "isoformat") else v) for k, v in spec.items()},
                    "call_ids": call_ids, "where": where, "param_keys": sorted(params), "join_calls": join})
    return out


def main():
    ref = ref_stub.load()
    rng = random.Random(20260210)
    chunks, artifacts = make_corpus(rng)
    out = []
    for case in cases(rng, chunks, artifacts):
        out.append({"case": compact(case), "response": run_reference(ref, case)})
    path = os.path.join(HERE, "reference_evidence.json")
    with open(path, "w") as f:
        json.dump({"settings": {"embeddings_exact_scan_threshold": ref.settings.embeddings_exact_scan_threshold,
                                "embeddings_hnsw_ef_search": ref.settings.embeddings_hnsw_ef_search},
                   "corpus": {"chunks": list(chunks.values()), "artifacts": list(artifacts.values())},
                   "resolve_call_ids": resolve_cases(ref), "filter_clause": filter_clause_cases(ref),
                   "cases": out}, f, default=str, separators=(",", ":"))
    print(f"wrote {path}: {len(out)} cases, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
