#!/usr/bin/env python
"""Generates tests/golden/reference_pure.json by running the REFERENCE's own pure functions
(imported live from /root/reference through oracle/ref_stub.py) on seeded inputs, and
tests/golden/dense_oracle.json from the C oracle (oracle/pgvector_restated.c) on the synthetic
corpus.  Run in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Floats are stored as float.hex() strings so that comparisons are bit-exact.
"""
import json
import os
import random
import sys
from datetime import datetime, timezone

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_stub  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def rrf_cases(ref):
    rng = random.Random(20260209)
    cases = []
    # the witness cases of SURVEY.md appendix A
    fixed = [
        {"bm25": [5, 7], "tech_tokens": [7, 9], "dense": [9, 5, 11]},
        {"bm25": [1], "tech_tokens": [1], "dense": [2, 1]},
        {"bm25": [], "tech_tokens": [], "dense": []},
        {"bm25": [3, 3, 3], "tech_tokens": [3]},            # duplicate ids inside one lane
        {"dense": list(range(100, 150))},
    ]
    for lanes in fixed:
        cases.append(lanes)
    for _ in range(60):
        n_lanes = rng.choice([1, 2, 3, 3, 3, 4])
        names = ["bm25", "tech_tokens", "dense", "extra"][:n_lanes]
        universe = rng.choice([20, 60, 200, 5000])
        lanes = {}
        for name in names:
            ln = rng.choice([0, 1, 5, 10, 50, 50])
            lanes[name] = rng.sample(range(1, universe + 1), min(ln, universe))
        cases.append(lanes)
    out = []
    for lanes in cases:
        for k in (60,) if len(out) % 7 else (60, 1, 1000):
            rows = {name: [{"chunk_id": i} for i in ids] for name, ids in lanes.items()}
            fused = ref.rrf_merge(rows, "chunk_id", k) if rows else []
            # lanes as an ORDERED list of [name, ids]: lane order decides the fp64 association
            out.append({"lanes": [[name, ids] for name, ids in lanes.items()], "k": k,
                        "fused": [[row["chunk_id"], sorted(hit), score.hex()] for row, hit, score in fused]})
    return out


def planner_cases(ref):
    F = ref.RetrieveFilters
    now = datetime(2026, 2, 9, tzinfo=timezone.utc)
    out = []
    saved = ref.settings.embeddings_exact_scan_threshold
    for threshold in (2000, 0, -5, 10, 5000):
        ref.settings.embeddings_exact_scan_threshold = threshold
        for rows in (-1, 0, 1, 9, 10, 11, 1999, 2000, 2001, 5000, 10**7):
            for kind in ("none", "nofilter_callids", "empty_callids", "date_from", "date_to", "tags",
                         "empty_tags", "filters_only_external"):
                filters, call_ids = None, None
                if kind == "nofilter_callids":
                    call_ids = ["c1"]
                elif kind == "empty_callids":
                    call_ids = []
                elif kind == "date_from":
                    filters = F(date_from=now)
                elif kind == "date_to":
                    filters = F(date_to=now)
                elif kind == "tags":
                    filters = F(call_tags=["x"])
                elif kind == "empty_tags":
                    filters = F(call_tags=[])
                elif kind == "filters_only_external":
                    filters = F(external_id="abc")
                out.append({"threshold": threshold, "rows": rows, "kind": kind,
                            "mode": ref.choose_dense_mode(rows, filters, call_ids),
                            "scoped": ref.dense_has_scoping(filters, call_ids)})
    ref.settings.embeddings_exact_scan_threshold = saved
    return out


def literal_cases(ref):
    import numpy as np
    rng = np.random.default_rng(7)
    vecs = [[0.1, 1 / 3, 1e-12, -0.0, 1.0], [float(np.float32(x)) for x in rng.standard_normal(64) / 32.0],
            [float(np.float32(x)) for x in (1e-38, 3.4e38, -1.17549435e-38, 0.0, 123456.789)]]
    return [{"values_hex": [float(v).hex() for v in vec], "literal": ref.vector_literal(vec)} for vec in vecs]


def token_cases(ref):
    texts = [
        "Which ticket tracked the ECONNRESET issue for ABC-123 on v1.2.3?",
        "We hit EAI_AGAIN and HTTP 503 from https://api.example.com/v2/items?id=7 at 10.1.2.3",
        "ORA-00942 after upgrading to 19.3; commit deadbeefcafe1234 touched /etc/app/config.yaml",
        "Bill of materials vs BOM: the build is building builds",
        "Lenovo and Dell vs. Super Micro (SMC) on AWS, Azure, GCP, OCI; Amazon Web Services; Google Cloud Platform",
        "object storage tiering with SSD; incumbent competitor; head to head bake-off; Oracle Cloud Infrastructure",
        "nothing technical here at all",
        "",
        "abc-123 ABC-123 Abc-123 JIRA-9 X-1 TOOLONGPREFIXX-12",
        "v1.2 1.2.3.4 1.2.3 v10.20.30 3.14",
        "E2BIG ENOENT Enoent E_ ECONNRESET econnreset",
        "/usr/local/bin/tool ./relative/path a/b/c /single",
        "Microsoft versus Google; competes competing competition competitive competitors compete",
        "HTTP500 http 404 Http 200 HTTPS 301",
        "1234567 abcdefg ABCDEF0 0123456789abcdef0123456789abcdef01234567 toolonghash0123456789abcdef0123456789abcdef0123456789",
    ]
    return [{"text": t, "tokens": ref.extract_tech_tokens(t)} for t in texts]


def debug_lane_cases(ref):
    rows = [{"chunk_id": 7, "score": 0.5, "text": "a"}, {"chunk_id": 3, "text": "b"}]
    return [{"rows": rows, "id_field": "chunk_id", "lane": ref.build_debug_lane(rows, "chunk_id")}]


def dense_oracle_cases():
    """Known answers of OUR restated oracle on the synthetic corpus (pins the oracle against
    regressions; it is not reference output -- the reference cannot run its dense lane here)."""
    import numpy as np
    from oracle import cpu_oracle as o
    out = []
    for n, k, nq in ((2000, 50, 4), (5000, 10, 2), (300, 50, 1)):
        x = o.synth_rows(20260209, 0, n)
        qs = o.synth_rows(20260210, 0, nq)
        for qi in range(nq):
            i64, s64 = o.exact_scan(qs[qi], x, k, variant=o.VARIANT_F64)
            i32, s32 = o.exact_scan(qs[qi], x, k, variant=o.VARIANT_PGV32)
            out.append({"n": n, "k": k, "query_row": qi, "ids_f64": i64.tolist(),
                        "scores_f64": [float(v).hex() for v in s64], "ids_pgv32": i32.tolist(),
                        "scores_pgv32": [float(v) for v in s32]})
    x = o.synth_rows(20260209, 0, 4)
    out_rows = {"first4_row_sums_hex": [float(np.float64(r.astype(np.float64).sum())).hex() for r in x],
                "row0_first8_hex": [float(v).hex() for v in x[0, :8]],
                "row3_last4_hex": [float(v).hex() for v in x[3, -4:]],
                "bf16_row0_first8": o.f32_to_bf16_bits(x[0, :8]).tolist(),
                "tag_bits_slot0_5": [int(o.lib().orc_synth_tag_bits(20260209, s)) for s in range(6)]}
    return {"scans": out, "generator": out_rows}


def main():
    ref = ref_stub.load()
    pure = {
        "source": "generated by tests/golden/make_golden.py from /root/reference (app/retrieve.py, app/ingest.py)",
        "rrf": rrf_cases(ref),
        "planner": planner_cases(ref),
        "vector_literal": literal_cases(ref),
        "tech_tokens": token_cases(ref),
        "debug_lane": debug_lane_cases(ref),
    }
    with open(os.path.join(HERE, "reference_pure.json"), "w") as f:
        json.dump(pure, f, indent=0, sort_keys=True)
    from oracle import cpu_oracle as o
    import ctypes
    o.lib().orc_synth_tag_bits.argtypes = [ctypes.c_uint64, ctypes.c_int64]
    o.lib().orc_synth_tag_bits.restype = ctypes.c_uint64
    with open(os.path.join(HERE, "dense_oracle.json"), "w") as f:
        json.dump(dense_oracle_cases(), f, indent=0, sort_keys=True)
    print("rrf", len(pure["rrf"]), "planner", len(pure["planner"]), "tokens", len(pure["tech_tokens"]))


if __name__ == "__main__":
    main()
