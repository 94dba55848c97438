"""tech_tokens lexical lane (host side) -- the second lane of the hybrid /retrieve.

  extract_tech_tokens(text)       app/ingest.py:141-160 (patterns app/ingest.py:24-73)
  TechTokenIndex.query(...)       the SQL of _fetch_chunks_tech, app/retrieve.py:195-208:
        WHERE <filters> AND tech_tokens && :tokens
        ORDER BY call_started_at DESC, chunk_id ASC LIMIT :limit

The reference evaluates ``tech_tokens && :tokens`` with a GIN index inside Postgres
(alembic/versions/0001_initial_schema.py:96); here it is a dictionary-encoded inverted index
(token -> sorted row list) on the host.  The lane produces ranks only (no score), so nothing in
it is floating point.  A device-resident version is SURVEY.md 8(f) row f-1.
"""
from __future__ import annotations

import re
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

# (regex, flags) -- order matters: findall results are emitted pattern by pattern (ingest.py:143-145)
_PATTERN_SPECS = (
    (r"https?://\S+", re.IGNORECASE),
    (r"\b(?:\d{1,3}\.){3}\d{1,3}\b", 0),              # IPv4
    (r"\b[A-Z]{2,10}-\d+\b", 0),                       # ticket ids
    (r"\bE[A-Z0-9_]{2,}\b", 0),                        # errno-style names
    (r"\bHTTP\s?\d{3}\b", re.IGNORECASE),
    (r"\bORA-\d{4,}\b", re.IGNORECASE),
    (r"\bv?\d+\.\d+(?:\.\d+)?\b", 0),                  # versions
    (r"\b[a-f0-9]{7,40}\b", re.IGNORECASE),            # commit hashes
    (r"(?:/[\w.\-]+)+", 0),                            # file paths
)
# canonical token <- trigger regex (all case-insensitive); order = emission order (ingest.py:146-148)
_DOMAIN_SPECS = (
    ("BOM", r"\bbill of materials\b"), ("BOM", r"\bbom\b"), ("build", r"\bbuild(?:s|ing)?\b"),
    ("SSD", r"\bssd\b"), ("object store", r"\bobject\s+(?:store|storage)\b"), ("object", r"\bobject\b"),
    ("tiering", r"\btiering\b"), ("Lenovo", r"\blenovo\b"), ("Dell", r"\bdell\b"),
    ("Supermicro", r"\bsuper[\s-]?micro\b|\bsmc\b"), ("AWS", r"\baws\b|\bamazon web services\b"),
    ("Amazon", r"\bamazon\b"), ("Azure", r"\bazure\b"), ("Microsoft", r"\bmicrosoft\b"),
    ("GCP", r"\bgcp\b|\bgoogle cloud(?: platform)?\b"), ("Google", r"\bgoogle\b"),
    ("OCI", r"\boci\b|\boracle cloud(?: infrastructure)?\b"), ("Oracle", r"\boracle\b"),
    ("competitive", r"\bcompet(?:e|es|ing|ition|itive|itor|itors)\b"), ("incumbent", r"\bincumbent\b"),
    ("bake-off", r"\bbake[\s-]?off\b"), ("head-to-head", r"\bhead[\s-]?to[\s-]?head\b"),
    ("vs", r"\bvs\.?(?=\s|$)|\bversus\b"),
)
_PATTERNS = tuple(re.compile(p, f) for p, f in _PATTERN_SPECS)
_DOMAIN = tuple((re.compile(p, re.IGNORECASE), canon) for canon, p in _DOMAIN_SPECS)
# one-shot prechecks: a text that matches none of the alternatives matches none of the patterns, so plain-language
# queries (the common case) cost two searches instead of thirty-two
_ANY_PATTERN = re.compile("|".join(f"(?i:{p})" if f & re.IGNORECASE else f"(?:{p})" for p, f in _PATTERN_SPECS))
_ANY_DOMAIN = re.compile("|".join(f"(?:{p})" for _canon, p in _DOMAIN_SPECS), re.IGNORECASE)


def extract_tech_tokens(text: str) -> List[str]:
    """Tokens of the exact-match lane: regex hits, then domain canonicals; stripped, empty dropped,
    first occurrence kept under case-insensitive comparison, original casing preserved."""
    found: List[str] = []
    if _ANY_PATTERN.search(text):
        for rx in _PATTERNS:
            found += rx.findall(text)
    if _ANY_DOMAIN.search(text):
        found += [canon for rx, canon in _DOMAIN if rx.search(text)]
    out: Dict[str, str] = {}
    for tok in found:
        tok = tok.strip()
        if tok:
            out.setdefault(tok.lower(), tok)
    return list(out.values())


class TechTokenIndex:
    """token -> ascending row list.  Element equality is case-sensitive, as text[] `&&` is."""

    def __init__(self):
        self._lists: Dict[str, List[int]] = {}
        self._arrays: Dict[str, np.ndarray] = {}
        self.version = 0            # bumped by every change: device copies know when they are out of date

    def add_row(self, row: int, tokens: Iterable[str]) -> None:
        for tok in set(tokens):
            self._lists.setdefault(tok, []).append(row)
        self._arrays.clear()
        self.version += 1

    def add_postings(self, token: str, rows: np.ndarray) -> None:
        """Bulk load (synthetic corpora): rows must be ascending."""
        self._arrays[token] = np.ascontiguousarray(rows, dtype=np.int64)
        self._lists.pop(token, None)
        self.version += 1

    def postings(self, token: str) -> np.ndarray:
        arr = self._arrays.get(token)
        if arr is None:
            arr = np.asarray(sorted(self._lists.get(token, ())), dtype=np.int64)
            self._arrays[token] = arr
        return arr

    def query(self, tokens: Sequence[str], cols: Dict[str, np.ndarray], limit: int, *,
              call_slots: Optional[Sequence[int]] = None, date_from=None, date_to=None,
              tag_mask: Optional[int] = None) -> np.ndarray:
        """Rows matching any token and the filter, ordered (call_started_at DESC, id ASC), first
        ``limit``.  ``cols`` are the store's host columns (ids, call_slot, started_at, tag_bits)."""
        from .store import to_micros
        lists = [self.postings(t) for t in tokens]
        lists = [l for l in lists if l.size]
        if not lists or limit <= 0:
            return np.empty(0, dtype=np.int64)
        rows = np.unique(np.concatenate(lists))
        keep = np.ones(rows.size, dtype=bool)
        if call_slots is not None:
            keep &= np.isin(cols["call_slot"][rows], np.asarray(list(call_slots), dtype=np.int32))
        if date_from is not None:
            keep &= cols["started_at"][rows] >= to_micros(date_from)
        if date_to is not None:
            keep &= cols["started_at"][rows] <= to_micros(date_to)
        if tag_mask is not None:
            keep &= (cols["tag_bits"][rows] & np.uint64(tag_mask)) != 0
        rows = rows[keep]
        order = np.lexsort((cols["ids"][rows], -cols["started_at"][rows]))
        return rows[order][:limit]


class DeviceTechIndex:
    """The same index resident in HBM: CSR postings over dictionary-encoded
    tokens + one precomputed rank per row; queries run in the ``tech_lane_kernel`` (csrc/tech_lane.cu)
    through ``cdr_tech_lane_host``.  Built once from a host :class:`TechTokenIndex` and the store's
    columns; a token unknown to the dictionary simply has no postings."""

    MAX_TOKENS = 32

    def __init__(self, host_index: TechTokenIndex, store):
        import ctypes
        from . import _ffi
        self._ffi = _ffi
        self.store = store
        self.host_index = host_index
        # what this copy was built from: a sealed store may grow and the host index with it (DenseStore.append /
        # TechTokenIndex.add_row); postings AND the global (call_started_at DESC, id ASC) ranks are then out of date
        self.built_rows = int(store.rows)
        self.built_version = int(host_index.version)
        tokens = sorted(set(host_index._lists) | set(host_index._arrays))
        self.token_ids: Dict[str, int] = {t: i for i, t in enumerate(tokens)}
        lists = [host_index.postings(t) for t in tokens]
        offsets = np.zeros(len(tokens) + 1, dtype=np.int64)
        if lists:
            offsets[1:] = np.cumsum([l.size for l in lists])
        rows = (np.concatenate(lists) if lists else np.empty(0, dtype=np.int64)).astype(np.uint32)
        cols = store.host_columns()
        order = np.lexsort((cols["ids"], -cols["started_at"]))          # (call_started_at DESC, id ASC)
        rank = np.empty(order.size, dtype=np.uint32)
        rank[order] = np.arange(order.size, dtype=np.uint32)
        self._h = ctypes.c_void_p()
        _ffi.check(_ffi.lib().cdr_tech_index_create(ctypes.byref(self._h), store.handle, _ffi.ptr(offsets),
                                                    len(tokens), _ffi.ptr(rows), _ffi.ptr(rank)),
                   "cdr_tech_index_create")

    def stale(self) -> bool:
        """True once the store or the host index changed after this copy was built."""
        return int(self.store.rows) != self.built_rows or int(self.host_index.version) != self.built_version

    def fits(self, tokens: Sequence[str]) -> bool:
        """A query's tokens fit the kernel's per-query token table (known tokens only count)."""
        return sum(1 for t in set(tokens) if t in self.token_ids) <= self.MAX_TOKENS

    def close(self) -> None:
        if self._h:
            self._ffi.lib().cdr_tech_index_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def encode_tokens(self, token_lists: Sequence[Sequence[str]]):
        """Token strings -> (ids int32 [nq, MAX_TOKENS] padded with -1, counts int32 [nq]); tokens the
        dictionary does not know have no postings and are dropped."""
        nq = len(token_lists)
        tok = np.full((nq, self.MAX_TOKENS), -1, dtype=np.int32)
        ntok = np.zeros(nq, dtype=np.int32)
        if self.stale():
            raise self._ffi.DenseEngineError("device tech-token index is out of date (the store or the host index changed "
                                             "after it was built): take it from DenseEngine.device_tech_index(), which "
                                             "rebuilds it", self._ffi.CDR_ERR_STATE)
        for i, toks in enumerate(token_lists):
            ids = [self.token_ids.get(t, -1) for t in dict.fromkeys(toks)]      # duplicates add nothing to `&&`
            ids = [t for t in ids if t >= 0]
            if len(ids) > self.MAX_TOKENS:
                raise self._ffi.DenseEngineError(f"query {i} carries {len(ids)} known tech tokens; the device lane takes "
                                                 f"{self.MAX_TOKENS} (callers route such requests to the host index)",
                                                 self._ffi.CDR_ERR_UNSUPPORTED)
            tok[i, :len(ids)] = ids
            ntok[i] = len(ids)
        return tok, ntok

    def query_batch(self, token_lists: Sequence[Sequence[str]], limit: int, *,
                    call_slots: Optional[Sequence[int]] = None, date_from=None, date_to=None,
                    tag_mask: Optional[int] = None):
        """ids[nq, limit] (unused slots -1) and n[nq] for nq token lists under one filter."""
        import ctypes
        import torch
        from .store import to_micros
        _ffi = self._ffi
        nq = len(token_lists)
        tok, ntok = self.encode_tokens(token_lists)
        bm, n_slots = self.store.slot_bitmap(call_slots)
        out_ids = np.empty((nq, limit), dtype=np.int64)
        out_n = np.empty(nq, dtype=np.int32)
        with torch.cuda.device(self.store.device):
            _ffi.check(_ffi.lib().cdr_tech_lane_host(
                self._h, _ffi.ptr(tok), _ffi.ptr(ntok), nq, self.MAX_TOKENS, _ffi.ptr(bm), n_slots,
                0 if date_from is None else 1, 0 if date_from is None else to_micros(date_from),
                0 if date_to is None else 1, 0 if date_to is None else to_micros(date_to),
                0 if tag_mask is None else 1, ctypes.c_uint64(tag_mask or 0), limit,
                _ffi.ptr(out_ids), _ffi.ptr(out_n), self.store._stream()), "cdr_tech_lane_host")
        return out_ids, out_n

    def query_ids(self, tokens: Sequence[str], limit: int, **spec) -> List[int]:
        if not tokens or limit <= 0:
            return []
        ids, n = self.query_batch([list(tokens)], limit, **spec)
        return ids[0, :int(n[0])].tolist()
