"""oracle/cpu_oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end to the C oracle (oracle/pgvector_restated.c, oracle/synth_ref.c) plus small
numpy restatements used to cross-check the C code itself.  Only tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module; the product package ``cadence_rag_b200`` never does.

PARITY UNPINNED (see the header of pgvector_restated.c and DESIGN.md): pgvector 0.8.1 is not
vendored in the reference and no reference test pins dense-lane results, so the dense oracle is
a restatement of pgvector's published algorithm anchored on the reference's SQL call sites
(app/retrieve.py:339-351, 374-386).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

VARIANT_PGV32 = 0   # fp32 accumulate, pgvector's loop ("what pgvector would say")
VARIANT_F64 = 1     # fp64 accumulate (ground-truth order)


def build(force: bool = False) -> str:
    """Compile the C oracle with oracle/Makefile (gcc only; seconds)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-s", "-C", _HERE, "_build/liboracle.so"], check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        c_fp = ctypes.c_void_p
        L.orc_synth_rows.argtypes = [c_fp, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        L.orc_synth_rows.restype = None
        L.orc_f32_to_bf16.argtypes = [c_fp, c_fp, ctypes.c_int64]
        L.orc_f32_to_bf16.restype = None
        L.orc_bf16_to_f32.argtypes = [c_fp, c_fp, ctypes.c_int64]
        L.orc_bf16_to_f32.restype = None
        L.orc_philox4x32_10.argtypes = [c_fp, c_fp, c_fp]
        L.orc_philox4x32_10.restype = None
        L.orc_cosine_distance_pgv32.argtypes = [c_fp, c_fp, ctypes.c_int]
        L.orc_cosine_distance_pgv32.restype = ctypes.c_double
        L.orc_cosine_distance_f64.argtypes = [c_fp, c_fp, ctypes.c_int]
        L.orc_cosine_distance_f64.restype = ctypes.c_double
        L.orc_exact_scan.argtypes = [c_fp, c_fp, c_fp, ctypes.c_int64, ctypes.c_int, c_fp,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_int, c_fp, c_fp]
        L.orc_exact_scan.restype = ctypes.c_int
        L.orc_all_scores.argtypes = [c_fp, c_fp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, c_fp]
        L.orc_all_scores.restype = None
        L.orc_exact_scan_bf16rows.argtypes = [c_fp, c_fp, c_fp, ctypes.c_int64, ctypes.c_int, c_fp,
                                              ctypes.c_int, c_fp, c_fp]
        L.orc_exact_scan_bf16rows.restype = ctypes.c_int
        L.orc_synth_tag_bits.argtypes = [ctypes.c_uint64, ctypes.c_int64]
        L.orc_synth_tag_bits.restype = ctypes.c_uint64
        L.orc_num_threads.argtypes = []
        L.orc_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


# ----------------------------------------------------------------------------- synthetic data
def synth_rows(seed: int, first_row: int, n: int, dim: int = 1024) -> np.ndarray:
    """fp32 [n, dim] rows of the synthetic corpus (global row ids first_row..first_row+n-1)."""
    assert dim % 4 == 0
    out = np.empty((n, dim), dtype=np.float32)
    lib().orc_synth_rows(_p(out), ctypes.c_uint64(seed), first_row, n, dim)
    return out


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.uint16)
    lib().orc_f32_to_bf16(_p(x), _p(out), x.size)
    return out


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(b, dtype=np.uint16)
    out = np.empty(b.shape, dtype=np.float32)
    lib().orc_bf16_to_f32(_p(b), _p(out), b.size)
    return out


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.empty(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def synth_rows_numpy(seed: int, first_row: int, n: int, dim: int = 1024) -> np.ndarray:
    """Pure-numpy restatement of synth_ref.c (cross-check of the C generator; small n only)."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = 0x9E3779B9, 0xBB67AE85
    rows = (np.arange(n, dtype=np.uint64) + np.uint64(first_row))[:, None]
    blk = np.arange(dim // 4, dtype=np.uint64)[None, :]
    c0 = np.broadcast_to(rows & np.uint64(0xFFFFFFFF), (n, dim // 4)).copy()
    c1 = np.broadcast_to(rows >> np.uint64(32), (n, dim // 4)).copy()
    c2 = np.broadcast_to(blk, (n, dim // 4)).copy()
    c3 = np.zeros((n, dim // 4), dtype=np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & mask
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & mask
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    w = np.stack([c0, c1, c2, c3], axis=-1).reshape(n, dim).astype(np.int64)
    s = (w & 0xFF) + ((w >> 8) & 0xFF) + ((w >> 16) & 0xFF) + (w >> 24) - 510
    sumsq = (s * s).sum(axis=1)
    with np.errstate(divide="ignore"):
        inv = np.where(sumsq > 0,
                       np.float32(1.0) / np.sqrt(sumsq.astype(np.float32)),
                       np.float32(0.0)).astype(np.float32)
    return (s.astype(np.float32) * inv[:, None]).astype(np.float32)


# ----------------------------------------------------------------------------- dense exact scan
def exact_scan(q: np.ndarray, x: np.ndarray, k: int, *, ids: Optional[np.ndarray] = None,
               allow: Optional[np.ndarray] = None, variant: int = VARIANT_F64,
               nthreads: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """The reference's dense SQL (app/retrieve.py:339-351) over an in-memory table.

    Returns (ids[m], scores[m]) with m <= k, ordered by distance ASC (NaN last), id ASC;
    score = 1 - distance.  ``allow`` is a uint32 bitmap (bit r => row r passes the WHERE clause
    incl. ``embedding IS NOT NULL``); ``ids`` default to row+1 (BIGSERIAL).
    """
    q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1)
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, dim = x.shape
    assert q.shape[0] == dim
    if ids is not None:
        ids = np.ascontiguousarray(ids, dtype=np.int64)
    if allow is not None:
        allow = np.ascontiguousarray(allow, dtype=np.uint32)
        assert allow.size * 32 >= n
    sc = np.empty(max(k, 1), dtype=np.float64)
    oi = np.empty(max(k, 1), dtype=np.int64)
    m = lib().orc_exact_scan(_p(q), _p(x), _p(ids), n, dim, _p(allow), k, variant, nthreads,
                             _p(sc), _p(oi))
    return oi[:m].copy(), sc[:m].copy()


def all_scores(q: np.ndarray, x: np.ndarray, variant: int = VARIANT_F64) -> np.ndarray:
    q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1)
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape[0], dtype=np.float64)
    lib().orc_all_scores(_p(q), _p(x), x.shape[0], x.shape[1], variant, _p(out))
    return out


def exact_scan_bf16rows(q: np.ndarray, xb_bits: np.ndarray, k: int, *,
                        ids: Optional[np.ndarray] = None,
                        allow: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1)
    xb = np.ascontiguousarray(xb_bits, dtype=np.uint16)
    n, dim = xb.shape
    if ids is not None:
        ids = np.ascontiguousarray(ids, dtype=np.int64)
    if allow is not None:
        allow = np.ascontiguousarray(allow, dtype=np.uint32)
    sc = np.empty(max(k, 1), dtype=np.float64)
    oi = np.empty(max(k, 1), dtype=np.int64)
    m = lib().orc_exact_scan_bf16rows(_p(q), _p(xb), _p(ids), n, dim, _p(allow), k, _p(sc), _p(oi))
    return oi[:m].copy(), sc[:m].copy()


def exact_scan_numpy_f64(q: np.ndarray, x: np.ndarray, k: int, *,
                         ids: Optional[np.ndarray] = None,
                         allow_rows: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Independent numpy fp64 restatement (cross-check of the C f64 variant; small inputs)."""
    q64 = np.asarray(q, dtype=np.float64).reshape(-1)
    x64 = np.asarray(x, dtype=np.float64)
    n = x64.shape[0]
    rid = np.arange(1, n + 1, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
    with np.errstate(invalid="ignore", divide="ignore"):
        sim = (x64 @ q64) / np.sqrt((x64 * x64).sum(axis=1) * (q64 * q64).sum())
    sim = np.where(sim > 1.0, 1.0, np.where(sim < -1.0, -1.0, sim))
    dist = 1.0 - sim
    keep = np.ones(n, dtype=bool) if allow_rows is None else np.asarray(allow_rows, dtype=bool)
    idx = np.nonzero(keep)[0]
    nan = np.isnan(dist[idx])
    order = np.lexsort((rid[idx], np.where(nan, 0.0, dist[idx]), nan))
    sel = idx[order][:k]
    return rid[sel].copy(), (1.0 - dist[sel]).copy()


def synth_tag_bits(seed: int, call_slot: int) -> int:
    return int(lib().orc_synth_tag_bits(ctypes.c_uint64(seed), call_slot))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def rows_to_bitmap(allow_rows: np.ndarray) -> np.ndarray:
    """bool[n] -> uint32 bitmap, bit (r & 31) of word (r >> 5)."""
    a = np.asarray(allow_rows, dtype=bool)
    n = a.shape[0]
    pad = (-n) % 32
    if pad:
        a = np.concatenate([a, np.zeros(pad, dtype=bool)])
    bits = np.packbits(a.reshape(-1, 32), axis=1, bitorder="little")
    return bits.view(np.uint32).reshape(-1).copy()


# ----------------------------------------------------------------------------- HNSW CPU baseline
class HnswBaseline:
    """pgvector's mode="ann" restated on the CPU (oracle/hnsw_baseline.cc: HNSW m=16,
    ef_construction=64, vector_cosine_ops) -- a BASELINE to time next to the GPU brute-force lane,
    "restated, not Postgres".  Rows are L2-normalised once at build time; the build is serial,
    queries run on `nthreads` OpenMP threads (one per query, like concurrent backends)."""

    _hlib = None

    @classmethod
    def _lib(cls):
        if cls._hlib is None:
            path = os.path.join(_HERE, "_build", "libhnsw_baseline.so")
            if not os.path.exists(path):
                subprocess.run(["make", "-s", "-C", _HERE, "_build/libhnsw_baseline.so"], check=True)
            L = ctypes.CDLL(path)
            vp = ctypes.c_void_p
            L.orc_hnsw_build.argtypes = [vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64]
            L.orc_hnsw_build.restype = vp
            L.orc_hnsw_search.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, vp, vp]
            L.orc_hnsw_search.restype = ctypes.c_int
            L.orc_hnsw_search_batch.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_int]
            L.orc_hnsw_search_batch.restype = None
            L.orc_hnsw_free.argtypes = [vp]
            L.orc_hnsw_free.restype = None
            cls._hlib = L
        return cls._hlib

    def __init__(self, x: np.ndarray, m: int = 16, ef_construction: int = 64, seed: int = 1):
        x = np.asarray(x, dtype=np.float32)
        norms = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
        self.x = np.ascontiguousarray((x / np.maximum(norms, 1e-30)).astype(np.float32))   # kept alive: the index points into it
        self.n, self.dim = self.x.shape
        self._h = self._lib().orc_hnsw_build(_p(self.x), self.n, self.dim, m, ef_construction, seed)

    def search(self, qs: np.ndarray, k: int, ef_search: int = 80, nthreads: int = 0):
        """rows[nq, k] (0-based, -1 = unused), sims[nq, k], n[nq]."""
        qs = np.asarray(qs, dtype=np.float32)
        if qs.ndim == 1:
            qs = qs[None, :]
        qn = np.ascontiguousarray((qs / np.maximum(np.linalg.norm(qs.astype(np.float64), axis=1, keepdims=True), 1e-30)).astype(np.float32))
        nq = qn.shape[0]
        rows = np.empty((nq, k), dtype=np.int64); sims = np.empty((nq, k), dtype=np.float32); n = np.empty(nq, dtype=np.int32)
        self._lib().orc_hnsw_search_batch(self._h, _p(qn), nq, k, ef_search, _p(rows), _p(sims), _p(n), nthreads)
        return rows, sims, n

    def close(self):
        if self._h:
            self._lib().orc_hnsw_free(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
