"""GPU tier: the batched bf16 tensor-core lane (K2, tcgen05) against the oracle.

Bar: recall@k >= 0.999 against the fp32 truth.  Because every survivor is re-scored
exactly (fp64 accumulate on the resident fp32 rows), the lane is expected to return the exact
lane's ids for all but boundary-noise cases; both are asserted.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from cadence_rag_b200 import _ffi  # noqa: E402
from cadence_rag_b200.store import DenseStore, SYNTH_CORPUS_SEED, SYNTH_QUERY_SEED  # noqa: E402
from oracle import cpu_oracle as orc  # noqa: E402


def _store(n, fp32=True, first_row=0):
    s = DenseStore("chunks", n, dim=1024, device=0, fp32=fp32, bf16=True)
    s.append_synthetic(n, first_row=first_row)
    s.finalize()
    return s


def _recall(got_ids, got_n, want_lists):
    hit = tot = 0
    for i, want in enumerate(want_lists):
        hit += len(set(want.tolist()) & set(got_ids[i, :int(got_n[i])].tolist()))
        tot += len(want)
    return hit / max(tot, 1)


@pytest.fixture(scope="module")
def corpus():
    n = 100_000
    s = _store(n)
    x = orc.synth_rows(SYNTH_CORPUS_SEED, 0, n)
    yield s, x
    s.close()


@pytest.mark.parametrize("nq,k", [(200, 50), (128, 10), (1, 50), (130, 100), (64, 192)])
def test_batch_lane_recall_and_exactness(corpus, nq, k):
    s, x = corpus
    qs = orc.synth_rows(SYNTH_QUERY_SEED, 1000, nq)
    ids, sc, cnt = s.search_batch(qs, k)
    assert np.all(cnt == k)
    e_ids, e_sc, e_cnt = s.search_exact(qs, k)
    check = min(nq, 24)
    want = [orc.exact_scan(qs[i], x, k, variant=orc.VARIANT_F64)[0] for i in range(check)]
    assert _recall(ids, cnt, want) >= 0.999
    # vs the exact lane over all queries
    same = sum(int(np.array_equal(ids[i], e_ids[i])) for i in range(nq))
    assert same >= int(np.floor(0.99 * nq)), f"{same}/{nq} queries identical to the exact lane"
    assert _recall(ids, cnt, [e_ids[i] for i in range(nq)]) >= 0.999
    for i in range(nq):
        if np.array_equal(ids[i], e_ids[i]):
            assert np.array_equal(sc[i].view(np.uint64), e_sc[i].view(np.uint64))
    assert np.all(np.diff(sc, axis=1) <= 0)
    # device-buffer entry point gives the same bits
    d_ids, d_sc, d_cnt = s.search_batch(torch.from_numpy(qs).cuda(), k)
    torch.cuda.synchronize()
    assert np.array_equal(d_ids.cpu().numpy(), ids)


def test_batch_lane_with_filter(corpus):
    s, x = corpus
    qs = orc.synth_rows(SYNTH_QUERY_SEED, 5000, 40)
    for spec in (dict(tag_mask=0b11), dict(call_slots=list(range(0, 500, 3))), dict(call_slots=[1, 2])):
        allow, count = s.filter_bitmap(**spec)
        ids, sc, cnt = s.search_batch(qs, 50, allow)
        e_ids, e_sc, e_cnt = s.search_exact(qs, 50, allow)
        assert np.array_equal(cnt, e_cnt)
        assert _recall(ids, cnt, [e_ids[i, :int(e_cnt[i])] for i in range(40)]) >= 0.999


@pytest.mark.parametrize("n", [1, 255, 256, 257, 4096, 5000, 70_000])
def test_batch_lane_ragged_sizes(n):
    s = _store(n)
    qs = orc.synth_rows(SYNTH_QUERY_SEED, 0, 9)
    ids, sc, cnt = s.search_batch(qs, 50)
    e_ids, e_sc, e_cnt = s.search_exact(qs, 50)
    assert np.array_equal(cnt, e_cnt) and np.all(cnt == min(n, 50))
    assert _recall(ids, cnt, [e_ids[i, :int(e_cnt[i])] for i in range(9)]) >= 0.999
    s.close()


def test_batch_lane_bf16_only_store():
    """C5 residency: bf16 rows only.  Truth = fp32 query x bf16-valued rows, fp64 accumulate."""
    n = 50_000
    s = _store(n, fp32=False)
    xb = s.read_rows(0, n, ("bf16",))["bf16"]
    qs = orc.synth_rows(SYNTH_QUERY_SEED, 77, 32)
    ids, sc, cnt = s.search_batch(qs, 50)
    want = [orc.exact_scan_bf16rows(qs[i], xb, 50) for i in range(32)]
    assert _recall(ids, cnt, [w[0] for w in want]) >= 0.999
    for i in range(32):
        if ids[i].tolist() == want[i][0].tolist():
            assert np.allclose(sc[i], want[i][1], rtol=1e-12)
    with pytest.raises(_ffi.DenseEngineError):
        s.search_exact(qs, 50)          # no fp32 rows resident -> loud error, no silent path
    s.close()


def test_batch_lane_thousands_of_identical_rows():
    """Adversarial corpus: 20 000 identical rows tie at the top.  Admission is decided on the packed key (score desc,
    row asc), so ties are cut by row order like in the exact lane and the result stays exact (ties by id)."""
    rng = np.random.default_rng(3)
    n = 30_000
    x = rng.standard_normal((n, 1024)).astype(np.float32)
    x[5000:25000] = x[5000]
    q = (x[5000] + 0.1 * rng.standard_normal(1024)).astype(np.float32)
    s = DenseStore("chunks", n, dim=1024, device=0)
    s.append(x, ids=np.arange(1, n + 1))
    s.finalize()
    qs = np.stack([q, rng.standard_normal(1024).astype(np.float32)])
    ids, sc, cnt = s.search_batch(qs, 50)
    for i in range(2):
        w_ids, w_sc = orc.exact_scan(qs[i], x, 50, variant=orc.VARIANT_F64)
        assert ids[i].tolist() == w_ids.tolist()
    s.close()


@pytest.mark.parametrize("fp32", [True, False])
def test_batch_lane_overflow_is_repaired_on_the_device(monkeypatch, fp32):
    """Candidate-list overflow (forced with a reduced list capacity: the first segment appends every row) is
    detected and repaired on the device: the overflowed queries are re-run on the exact lane (fp32 stores) or the
    bf16 scan lane (bf16-only stores) by kernels that read the query list from device memory -- the call never
    synchronises.  Results must carry that lane's bits; queries that did not overflow keep the tensor-core result."""
    n = 20_000
    s = _store(n, fp32=fp32)
    qs = torch.from_numpy(orc.synth_rows(SYNTH_QUERY_SEED, 300, 1100)).cuda()        # 1100 queries: two re-run rounds
    exact = (s.search_exact if fp32 else s.search_scan_bf16)(qs, 50)
    plain = s.search_batch(qs, 50)
    monkeypatch.setenv("CADENCE_K2_TEST_CAP", "2000")                                # < 4096 rows of the first segment
    redo = s.search_batch(qs, 50)
    torch.cuda.synchronize()
    for a, b in zip(redo, exact):
        assert torch.equal(a.view(torch.int64) if a.dtype == torch.float64 else a, b.view(torch.int64) if b.dtype == torch.float64 else b)
    monkeypatch.delenv("CADENCE_K2_TEST_CAP")
    again = s.search_batch(qs, 50)
    torch.cuda.synchronize()
    assert torch.equal(again[0], plain[0])
    rec = _recall(plain[0].cpu().numpy(), plain[2].cpu().numpy(), [r for r in exact[0].cpu().numpy()])
    assert rec >= 0.999
    s.close()


def test_batch_lane_nan_rows_and_tight_filters_need_no_second_pass():
    """Rows whose score is NaN (zero embeddings) are admitted by the tensor-core lane exactly as by the exact lane
    (below every real score, id order), so short and NaN-tailed results agree with the exact lane without any
    fallback -- on bf16-only stores too, where no exact fp32 lane exists."""
    rng = np.random.default_rng(11)
    n = 6_000
    x = rng.standard_normal((n, 1024)).astype(np.float32)
    x[100:140] = 0.0                                            # 40 zero rows: NaN cosine
    qs = rng.standard_normal((5, 1024)).astype(np.float32)
    for fp32 in (True, False):
        s = DenseStore("chunks", n, dim=1024, device=0, fp32=fp32, bf16=True)
        s.append(x, ids=np.arange(1, n + 1), call_ids=[r // 20 for r in range(n)])
        s.finalize()
        other = s.search_exact if fp32 else s.search_scan_bf16
        for slots in ([5, 6], [4, 5, 6], [0, 1], list(range(300))):   # 40 rows incl. 20 NaN / 60 rows / 40 real rows / all
            allow, count = s.filter_bitmap(call_slots=slots)
            ids, sc, cnt = s.search_batch(qs, 50, allow)
            e_ids, e_sc, e_cnt = other(qs, 50, allow)
            assert np.array_equal(cnt, e_cnt) and np.all(cnt == min(count, 50))
            if count <= 128:                                     # every allowed row is a candidate: identical lists
                assert np.array_equal(ids, e_ids)
                assert np.array_equal(np.isnan(sc), np.isnan(e_sc))
            else:
                assert _recall(ids, cnt, [e_ids[i, :int(e_cnt[i])] for i in range(5)]) >= 0.999
        s.close()


def test_batch_lane_vs_torch_fp32_matmul_reference_1m():
    """An oracle that shares no code with the engine: plain PyTorch fp32 matmul (TF32 off) over the regenerated
    1 M x 1024 corpus in row chunks with a running top-k, for all 1024 queries of a batch.
    The tensor-core lane must reach recall@50 >= 0.999 against it, and where the ids agree its fp64 scores must
    equal the fp32 reference within 1e-5 relative; the exact fp32 lane is held to the same reference."""
    from cadence_rag_b200.store import synth_rows_device
    n, nq, k = 1_000_000, 1024, 50
    free, _ = torch.cuda.mem_get_info()
    if free < 20e9:
        pytest.skip("needs ~12 GB of free HBM")
    s = _store(n)
    old_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        q = synth_rows_device(SYNTH_QUERY_SEED, 5000, nq, 1024, device=0)
        qn = q / q.norm(dim=1, keepdim=True)
        best_sc = torch.full((nq, k), -2.0, device="cuda")
        best_id = torch.full((nq, k), -1, dtype=torch.int64, device="cuda")
        chunk = 100_000
        for r0 in range(0, n, chunk):
            x = synth_rows_device(SYNTH_CORPUS_SEED, r0, chunk, 1024, device=0)      # same bits as the store's rows
            sc = qn @ (x / x.norm(dim=1, keepdim=True)).T                             # fp32 cosine, [nq, chunk]
            c_sc, c_ix = sc.topk(k, dim=1)
            all_sc = torch.cat([best_sc, c_sc], dim=1)
            all_id = torch.cat([best_id, c_ix + (r0 + 1)], dim=1)                     # chunk_id = row + 1
            best_sc, pick = all_sc.topk(k, dim=1)
            best_id = all_id.gather(1, pick)
        ref_id, ref_sc = best_id.cpu().numpy(), best_sc.cpu().numpy().astype(np.float64)
        for name, (ids, sc, cnt) in (("bf16 tensor-core lane", s.search_batch(q, k)), ("exact fp32 lane", s.search_exact(q[:128].contiguous(), k))):
            torch.cuda.synchronize()
            ids, sc, cnt = ids.cpu().numpy(), sc.cpu().numpy(), cnt.cpu().numpy()
            m = ids.shape[0]
            assert np.all(cnt == k), name
            recall = np.mean([len(set(ids[i]) & set(ref_id[i])) / k for i in range(m)])
            assert recall >= 0.999, (name, recall)
            agree = ids == ref_id[:m]
            assert agree.mean() > 0.97, (name, agree.mean())                           # fp32 vs fp64 order: rare near-tie swaps only
            assert np.allclose(sc[agree], ref_sc[:m][agree], rtol=1e-5, atol=0), name
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old_tf32
        s.close()


def test_batch_lane_skips_empty_tiles_under_block_filters(corpus):
    """Filters with block structure (a date range over time-ordered rows, a run of calls, a single row) leave whole
    256-row corpus tiles without an allowed row; the tensor-core lane skips those tiles in all three warp roles.
    The answer must still be the exact lane's."""
    from cadence_rag_b200.store import SYNTH_CALL_PERIOD_US, SYNTH_T0_US
    s, x = corpus
    n = x.shape[0]
    qs = orc.synth_rows(SYNTH_QUERY_SEED, 7000, 130)
    singles = np.zeros(n, dtype=bool); singles[[5, 255, 256, 70_000, n - 1]] = True
    specs = [dict(date_from=SYNTH_T0_US + 300 * SYNTH_CALL_PERIOD_US),                        # the last 40 % of the rows
             dict(date_from=SYNTH_T0_US + 100 * SYNTH_CALL_PERIOD_US, date_to=SYNTH_T0_US + 130 * SYNTH_CALL_PERIOD_US),
             dict(call_slots=list(range(40, 60)) + [499])]
    allows = [s.filter_bitmap(**sp)[0] for sp in specs]
    allows.append(torch.from_numpy(orc.rows_to_bitmap(singles).view(np.int32)).cuda())       # 5 rows: short result -> exact-lane fallback
    for allow in allows:
        for nq in (130, 256):
            q = np.concatenate([qs, qs[:nq - 130]]) if nq > 130 else qs
            ids, sc, cnt = s.search_batch(q, 50, allow)
            e_ids, e_sc, e_cnt = s.search_exact(q, 50, allow)
            assert np.array_equal(cnt, e_cnt)
            assert _recall(ids, cnt, [e_ids[i, :int(e_cnt[i])] for i in range(nq)]) >= 0.999
            same = ids == e_ids
            assert same.mean() > 0.98 and np.array_equal(sc[same & (ids >= 0)].view(np.uint64), e_sc[same & (ids >= 0)].view(np.uint64))
