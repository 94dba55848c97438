"""CPU tier: the oracle is pinned before it is trusted.

* Philox core against the Random123 known-answer vectors.
* C generator == independent numpy restatement (bit-exact).
* C dense oracle (pgv32 / f64 variants) == numpy fp64 restatement, == committed known answers.
* Pure-Python ports of the reference's pure functions == golden vectors produced by the
  reference's own code (tests/golden/reference_pure.json), and == the live reference when
  /root/reference is present.
"""
import json
import os
import types

import numpy as np
import pytest

from oracle import cpu_oracle as orc
from oracle import ports, ref_stub


@pytest.fixture(scope="module")
def pure(golden_dir):
    with open(os.path.join(golden_dir, "reference_pure.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def dense(golden_dir):
    with open(os.path.join(golden_dir, "dense_oracle.json")) as f:
        return json.load(f)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert orc.philox4x32_10(ctr, key).tolist() == want


@pytest.mark.parametrize("dim,first,n", [(1024, 0, 33), (1024, 2**33 + 5, 7), (256, 12345, 16), (8, 0, 4)])
def test_generator_c_equals_numpy(dim, first, n):
    a = orc.synth_rows(20260209, first, n, dim)
    b = orc.synth_rows_numpy(20260209, first, n, dim)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    if dim >= 256:
        norms = np.linalg.norm(a.astype(np.float64), axis=1)
        assert np.all(np.abs(norms - 1.0) < 1e-6)


def test_generator_is_shard_invariant():
    whole = orc.synth_rows(7, 100, 50)
    parts = np.concatenate([orc.synth_rows(7, 100, 20), orc.synth_rows(7, 120, 30)])
    assert np.array_equal(whole.view(np.uint32), parts.view(np.uint32))


def test_generator_known_answers(dense):
    g = dense["generator"]
    x = orc.synth_rows(20260209, 0, 4)
    assert [float(v).hex() for v in x[0, :8]] == g["row0_first8_hex"]
    assert [float(v).hex() for v in x[3, -4:]] == g["row3_last4_hex"]
    assert orc.f32_to_bf16_bits(x[0, :8]).tolist() == g["bf16_row0_first8"]
    assert [orc.synth_tag_bits(20260209, s) for s in range(6)] == g["tag_bits_slot0_5"]


def test_bf16_rounding_matches_torch():
    torch = pytest.importorskip("torch")
    x = orc.synth_rows(3, 0, 8)
    x[0, :4] = [1.0 + 2**-8, 1.0 + 3 * 2**-8, -0.0, 65504.0]   # ties-to-even cases
    want = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(orc.f32_to_bf16_bits(x), want)
    back = orc.bf16_bits_to_f32(want)
    assert np.array_equal(back, torch.from_numpy(x).to(torch.bfloat16).float().numpy())


def test_dense_oracle_known_answers(dense):
    for case in dense["scans"]:
        x = orc.synth_rows(20260209, 0, case["n"])
        q = orc.synth_rows(20260210, case["query_row"], 1)[0]
        ids, sc = orc.exact_scan(q, x, case["k"], variant=orc.VARIANT_F64)
        assert ids.tolist() == case["ids_f64"]
        assert [float(v).hex() for v in sc] == case["scores_f64"]
        ids32, sc32 = orc.exact_scan(q, x, case["k"], variant=orc.VARIANT_PGV32)
        # the fp32 loop's summation order is compiler-defined: ids must agree, scores to 1e-6
        assert ids32.tolist() == case["ids_pgv32"]
        assert np.allclose(sc32, case["scores_pgv32"], rtol=0, atol=1e-6)


def test_dense_oracle_c_equals_numpy():
    x = orc.synth_rows(11, 0, 3000)
    q = orc.synth_rows(12, 5, 1)[0]
    allow_rows = (np.arange(3000) % 3) != 1
    ids_c, sc_c = orc.exact_scan(q, x, 50, allow=orc.rows_to_bitmap(allow_rows), variant=orc.VARIANT_F64)
    ids_n, sc_n = orc.exact_scan_numpy_f64(q, x, 50, allow_rows=allow_rows)
    assert ids_c.tolist() == ids_n.tolist()
    assert np.allclose(sc_c, sc_n, rtol=0, atol=1e-14)
    ids32, sc32 = orc.exact_scan(q, x, 50, allow=orc.rows_to_bitmap(allow_rows), variant=orc.VARIANT_PGV32)
    assert ids32.tolist() == ids_n.tolist()
    assert np.allclose(sc32, sc_n, rtol=1e-5, atol=0)
    # single thread == all threads
    ids1, sc1 = orc.exact_scan(q, x, 50, variant=orc.VARIANT_PGV32, nthreads=1)
    idsN, scN = orc.exact_scan(q, x, 50, variant=orc.VARIANT_PGV32, nthreads=0)
    assert ids1.tolist() == idsN.tolist() and np.array_equal(sc1, scN)


def test_dense_oracle_edge_semantics():
    """Known-answer micro fixtures."""
    dim = 16
    x = np.zeros((6, dim), dtype=np.float32)
    x[0, 0] = 1.0          # e0
    x[1, 1] = 2.0          # e1 scaled
    x[2, 0] = 3.0          # duplicate direction of row 0 -> tie, id order
    x[3] = 0.0             # zero vector -> NaN distance -> last
    x[4, 0] = -1.0         # opposite
    x[5, 0] = 1.0; x[5, 1] = 1.0
    q = np.zeros(dim, dtype=np.float32); q[0] = 5.0
    for variant in (orc.VARIANT_PGV32, orc.VARIANT_F64):
        ids, sc = orc.exact_scan(q, x, 10, variant=variant)
        assert ids.tolist() == [1, 3, 6, 2, 5, 4]
        assert sc[0] == 1.0 and sc[1] == 1.0 and sc[3] == 0.0 and sc[4] == -1.0
        assert abs(sc[2] - 2 ** -0.5) < 1e-7 and np.isnan(sc[5])
        # LIMIT larger than survivors -> short list; filter excludes rows
        allow = orc.rows_to_bitmap(np.array([1, 0, 1, 0, 0, 0], dtype=bool))
        ids, sc = orc.exact_scan(q, x, 10, allow=allow, variant=variant)
        assert ids.tolist() == [1, 3]
        ids, sc = orc.exact_scan(q, x, 1, variant=variant)
        assert ids.tolist() == [1]
        # custom ids
        ids, _ = orc.exact_scan(q, x, 2, ids=np.array([10, 20, 30, 40, 50, 60]), variant=variant)
        assert ids.tolist() == [10, 30]


# ------------------------------------------------------------------ ports vs reference goldens
def _rows(ids):
    return [{"chunk_id": i} for i in ids]


def test_port_rrf_matches_reference_golden(pure):
    for case in pure["rrf"]:
        lanes = {name: _rows(ids) for name, ids in case["lanes"]}
        fused = ports.rrf_merge(lanes, "chunk_id", case["k"]) if lanes else []
        got = [[row["chunk_id"], sorted(hit), score.hex()] for row, hit, score in fused]
        assert got == case["fused"]


def test_rrf_association_witness():
    # (1/61 + 1/61) + 1/62 differs from the other association at 1 ulp
    a = (0.0 + 1.0 / 61) + 1.0 / 61
    a = a + 1.0 / 62
    assert a == 0.04891591750396616
    b = (0.0 + 1.0 / 62) + 1.0 / 61
    b = b + 1.0 / 61
    assert b == 0.048915917503966164 and a != b


def test_port_planner_matches_reference_golden(pure):
    for case in pure["planner"]:
        F = types.SimpleNamespace
        kind = case["kind"]
        filters, call_ids = None, None
        if kind == "nofilter_callids":
            call_ids = ["c1"]
        elif kind == "empty_callids":
            call_ids = []
        elif kind == "date_from":
            filters = F(date_from=1, date_to=None, call_tags=None)
        elif kind == "date_to":
            filters = F(date_from=None, date_to=1, call_tags=None)
        elif kind == "tags":
            filters = F(date_from=None, date_to=None, call_tags=["x"])
        elif kind == "empty_tags":
            filters = F(date_from=None, date_to=None, call_tags=[])
        elif kind == "filters_only_external":
            filters = F(date_from=None, date_to=None, call_tags=None, external_id="abc")
        assert ports.choose_dense_mode(case["rows"], filters, call_ids, case["threshold"]) == case["mode"]
        assert ports.dense_has_scoping(filters, call_ids) == case["scoped"]


def test_port_vector_literal_matches_reference_golden(pure):
    for case in pure["vector_literal"]:
        vals = [float.fromhex(h) for h in case["values_hex"]]
        assert ports.vector_literal(vals) == case["literal"]
        # .10g round-trips float32 exactly
        f32 = np.array(vals, dtype=np.float32)
        lit = ports.vector_literal(f32.tolist())
        back = np.array([np.float32(t) for t in lit[1:-1].split(",")], dtype=np.float32)
        assert np.array_equal(back.view(np.uint32), f32.view(np.uint32))


@pytest.mark.skipif(not ref_stub.available(), reason="/root/reference not present (GPU box)")
def test_ports_match_live_reference():
    ref = ref_stub.load()
    rng = np.random.default_rng(5)
    for _ in range(50):
        lanes = {}
        for name in ("bm25", "tech_tokens", "dense"):
            n = int(rng.integers(0, 51))
            lanes[name] = _rows(rng.choice(np.arange(1, 120), size=n, replace=False).tolist())
        want = ref.rrf_merge(lanes, "chunk_id")
        got = ports.rrf_merge(lanes, "chunk_id")
        assert [(r["chunk_id"], h, s) for r, h, s in got] == [(r["chunk_id"], h, s) for r, h, s in want]
    vals = rng.standard_normal(100).tolist()
    assert ports.vector_literal(vals) == ref.vector_literal(vals)


def test_hnsw_baseline_restatement_sane():
    """The CPU HNSW baseline (bench.py's cpu_baseline_hnsw; pgvector's m=16 / ef_construction=64 / ef_search=80)
    finds the exact neighbours on data with low intrinsic dimension and is independent of the thread count."""
    rng = np.random.default_rng(3)
    n, k = 1500, 20
    z = rng.standard_normal((n, 12)).astype(np.float32)
    proj = rng.standard_normal((12, 256)).astype(np.float32)
    x = z @ proj
    qs = rng.standard_normal((16, 12)).astype(np.float32) @ proj
    index = orc.HnswBaseline(x)
    rows, sims, cnt = index.search(qs, k, 80)
    rows1, _, _ = index.search(qs, k, 80, nthreads=1)
    assert np.array_equal(rows, rows1) and (cnt == k).all()
    hit = 0
    for i in range(qs.shape[0]):
        want, want_sc = orc.exact_scan(qs[i], x, k)
        hit += len(set(rows[i].tolist()) & set((want - 1).tolist()))
        assert np.all(np.diff(sims[i]) <= 1e-6)                      # closest first
        assert abs(float(sims[i][0]) - float(want_sc[0])) < 1e-4     # cosine of the best hit
    assert hit / (k * qs.shape[0]) >= 0.97
    index.close()
