"""Resident corpus store -- host-side owner of one ``cdr_store`` (one per reference table).

Replaces what the reference keeps in Postgres for the dense lane: the ``embedding vector(1024)``
column and the filter columns of ``chunks`` / ``artifact_chunks``
(alembic/versions/0001_initial_schema.py:78-87, 0006_add_artifact_chunks.py:22-33), plus the
dictionaries that turn call UUIDs and tag strings into the codes the device columns hold.
"""
from __future__ import annotations

import ctypes
from datetime import datetime, timezone
from typing import Any, Dict, Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np

from . import _ffi
from ._ffi import DenseEngineError

_EPOCH = datetime(1970, 1, 1, tzinfo=timezone.utc)

SYNTH_CORPUS_SEED = 20260209     # BASELINE.md "Data"
SYNTH_QUERY_SEED = 20260210
SYNTH_ROWS_PER_CALL = 200
SYNTH_T0_US = 1_700_000_000_000_000
SYNTH_CALL_PERIOD_US = 3_600_000_000


def to_micros(value) -> int:
    """datetime (naive => UTC, like a timestamptz session in UTC) or int microseconds -> int."""
    if isinstance(value, datetime):
        if value.tzinfo is None:
            value = value.replace(tzinfo=timezone.utc)
        delta = value - _EPOCH
        return (delta.days * 86400 + delta.seconds) * 1_000_000 + delta.microseconds
    return int(value)


def _encode_call_id(c):
    """JSON form of a call id with its type kept (snapshots): UUID, int or str."""
    import uuid
    if isinstance(c, uuid.UUID):
        return ["uuid", str(c)]
    if isinstance(c, (int, np.integer)):
        return ["int", int(c)]
    return ["str", str(c)]


def _decode_call_id(c):
    import uuid
    kind, value = c
    return uuid.UUID(value) if kind == "uuid" else (int(value) if kind == "int" else value)


def _torch():
    import torch
    return torch


def synth_rows_device(seed: int, first_row: int, n: int, dim: int, device: Optional[int] = None):
    """[n, dim] fp32 CUDA tensor of synthetic rows (global rows first_row..) -- used for queries."""
    torch = _torch()
    _ffi.require_device()
    from .config import settings
    dev = settings.cadence_gpu_device if device is None else device
    with torch.cuda.device(dev):
        out = torch.empty((n, dim), dtype=torch.float32, device=f"cuda:{dev}")
        _ffi.check(_ffi.lib().cdr_synth_rows(_ffi.ptr(out), ctypes.c_uint64(seed), first_row, n, dim,
                                             _ffi.stream_ptr()), "cdr_synth_rows")
    return out


class DenseStore:
    """One table's resident embeddings + filter columns on one GPU."""

    def __init__(self, table_name: str, capacity_rows: int, dim: Optional[int] = None,
                 device: Optional[int] = None, fp32: bool = True, bf16: bool = True,
                 key_field: Optional[str] = None):
        from .config import settings
        _ffi.require_device()
        self.table_name = table_name
        self.key_field = key_field or ("artifact_chunk_id" if table_name == "artifact_chunks" else "chunk_id")
        self.dim = int(dim or max(1, int(settings.embeddings_dim)))
        self.device = settings.cadence_gpu_device if device is None else int(device)
        self.flags = (_ffi.CDR_STORE_FP32 if fp32 else 0) | (_ffi.CDR_STORE_BF16 if bf16 else 0)
        self.capacity = int(capacity_rows)
        self._h = ctypes.c_void_p()
        _ffi.check(_ffi.lib().cdr_store_create(ctypes.byref(self._h), self.device, self.capacity,
                                               self.dim, self.flags), "cdr_store_create")
        self.finalized = False
        # dictionaries (host): call UUID -> slot, tag -> bit
        self.call_slots: Dict[Any, int] = {}
        self.call_ids_by_slot: List[Any] = []
        self.tag_bits: Dict[str, int] = {}
        # tags beyond the 64 bits of the device column ("overflow -> host-side call set"):
        # tag -> the call slots that carry it; a filter naming such a tag is served as a call-slot set
        self.overflow_tag_slots: Dict[str, Set[int]] = {}
        self.slot_tag_mask: Dict[int, int] = {}       # call slot -> its hot-tag mask (the calls.tags of that call)
        # optional payload columns returned with each hit (speaker, text, ... as in the SQL SELECT)
        self.payload: Dict[int, Dict[str, Any]] = {}
        self.synthetic = None   # (seed, first_row) when filled by the on-device generator
        self._host_cols: Optional[Dict[str, np.ndarray]] = None
        self._rows: Optional[int] = None

    # ------------------------------------------------------------------ lifecycle
    @property
    def handle(self) -> ctypes.c_void_p:
        if not self._h:
            raise DenseEngineError("store destroyed", _ffi.CDR_ERR_STATE)
        return self._h

    def close(self) -> None:
        if self._h:
            _ffi.lib().cdr_store_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return _ffi.current_stream_ptr(self.device)

    # ------------------------------------------------------------------ dictionaries
    def slot_of_call(self, call_id, create: bool = False) -> Optional[int]:
        slot = self.call_slots.get(call_id)
        if slot is None and create:
            slot = len(self.call_ids_by_slot)
            self.call_slots[call_id] = slot
            self.call_ids_by_slot.append(call_id)
        return slot

    def bits_of_tags(self, tags: Optional[Iterable[str]], create: bool = False, slot: Optional[int] = None) -> int:
        """Hot-tag mask of a tag list.  ``create`` (ingest): unknown tags take the next free bit; once the 64 bits of
        the device column are taken, further tags are kept host-side as tag -> call-slot sets (``slot`` = the call
        the row belongs to: tags are a property of the call, ``calls.tags`` in the reference's schema)."""
        mask = 0
        for tag in tags or ():
            bit = self.tag_bits.get(tag)
            if bit is None and create and tag not in self.overflow_tag_slots and len(self.tag_bits) < 64:
                bit = len(self.tag_bits)
                self.tag_bits[tag] = bit
            if bit is not None:
                mask |= 1 << bit
            elif create:
                if slot is None:
                    raise DenseEngineError(f"tag {tag!r} is beyond the 64 device tag bits and needs the row's call "
                                           "(pass call_ids with call_tags)", _ffi.CDR_ERR_INVALID)
                self.overflow_tag_slots.setdefault(tag, set()).add(int(slot))
        if create and slot is not None and mask:
            self.slot_tag_mask[int(slot)] = self.slot_tag_mask.get(int(slot), 0) | mask
        return mask

    def tag_filter(self, tags: Iterable[str]):
        """``c.tags && :call_tags`` (app/retrieve.py:112-115) as store codes: (tag_mask, None) when every known
        tag of the filter has a device bit; (None, sorted call slots) when the filter names an overflow tag -- the
        calls carrying ANY of the tags, hot ones included, as one call-slot set."""
        tags = list(tags)
        mask = self.bits_of_tags(tags)
        extra = [self.overflow_tag_slots[t] for t in tags if t in self.overflow_tag_slots]
        if not extra:
            return mask, None
        slots = set().union(*extra)
        if mask:
            slots.update(sl for sl, m in self.slot_tag_mask.items() if m & mask)
        return None, sorted(slots)

    # ------------------------------------------------------------------ ingest
    def append(self, embeddings, ids: Sequence[int], call_ids: Optional[Sequence[Any]] = None,
               call_started_at: Optional[Sequence[Any]] = None,
               call_tags: Optional[Sequence[Optional[Iterable[str]]]] = None,
               valid: Optional[Sequence[bool]] = None,
               payload: Optional[Sequence[Dict[str, Any]]] = None) -> None:
        """Append rows (the counterpart of embedding_pipeline._update_embeddings,
        app/embedding_pipeline.py:149-168).  ``embeddings``: [n, dim] float32 (numpy, torch CPU or
        torch CUDA); rows with ``valid[i] == False`` model ``embedding IS NULL``."""
        torch = _torch()
        n = len(ids)
        if n == 0:
            return
        is_dev = hasattr(embeddings, "is_cuda") and embeddings.is_cuda
        if is_dev:
            emb = embeddings.to(dtype=torch.float32).contiguous()
        else:
            emb = np.ascontiguousarray(np.asarray(embeddings, dtype=np.float32))
        if tuple(emb.shape) != (n, self.dim):
            raise DenseEngineError(f"embeddings shape {tuple(emb.shape)} != ({n}, {self.dim})")
        ids_np = np.ascontiguousarray(np.asarray(ids, dtype=np.int64))
        slots = np.zeros(n, dtype=np.int32)
        if call_ids is not None:
            slots = np.fromiter((self.slot_of_call(c, create=True) for c in call_ids), dtype=np.int32, count=n)
        started = np.zeros(n, dtype=np.int64)
        if call_started_at is not None:
            started = np.fromiter((to_micros(t) for t in call_started_at), dtype=np.int64, count=n)
        tags = np.zeros(n, dtype=np.uint64)
        if call_tags is not None:
            tags = np.fromiter((self.bits_of_tags(t, create=True, slot=int(slots[i]) if call_ids is not None else None)
                                for i, t in enumerate(call_tags)), dtype=np.uint64, count=n)
        valid_np = None
        if valid is not None:
            valid_np = np.ascontiguousarray(np.asarray(valid, dtype=np.uint8))

        def dev(a):
            return torch.from_numpy(a).to(f"cuda:{self.device}") if is_dev else a

        cols = [dev(ids_np), dev(slots), dev(started), dev(tags.view(np.int64)), None if valid_np is None else dev(valid_np)]
        with torch.cuda.device(self.device):
            _ffi.check(_ffi.lib().cdr_store_append(
                self.handle, _ffi.ptr(emb), _ffi.ptr(cols[0]), _ffi.ptr(cols[1]), _ffi.ptr(cols[2]),
                _ffi.ptr(cols[3]), _ffi.ptr(cols[4]), n, 1 if is_dev else 0, self._stream()), "cdr_store_append")
            if is_dev:
                torch.cuda.current_stream(self.device).synchronize()
        if payload is not None:
            for i, row in zip(ids_np.tolist(), payload):
                self.payload[i] = row
        self._host_cols = None      # a sealed store may grow: drop the cached host columns / row count
        self._rows = None

    def update_embeddings(self, ids: Sequence[int], vectors) -> None:
        """``UPDATE <table> SET embedding = CAST(:embedding AS vector(D)) WHERE <id> = :row_id`` for rows that
        already exist (app/embedding_pipeline.py:149-168, _update_embeddings) -- the backfill of rows ingested
        with ``embedding IS NULL``.  ``vectors``: [n, dim] floats, or the reference's ``"[...]"`` literals
        (parsed with float32 rounding like pgvector's vector_in).  All-or-nothing: an unknown id raises and
        nothing is updated."""
        n = len(ids)
        if n == 0:
            return
        if len(vectors) != n:
            raise DenseEngineError(f"row/vector mismatch for {self.table_name}: {n} rows vs {len(vectors)} vectors")
        if isinstance(vectors[0], str):
            emb = np.zeros((n, self.dim), dtype=np.float32)
            for i, lit in enumerate(vectors):
                body = lit.strip()
                if not (body.startswith("[") and body.endswith("]")):
                    raise DenseEngineError(f"row {i}: malformed vector literal")
                vals = np.array(body[1:-1].split(","), dtype=np.float32)
                if vals.shape[0] != self.dim:
                    raise DenseEngineError(f"row {i}: expected {self.dim} dimensions, not {vals.shape[0]}")
                emb[i] = vals
        else:
            emb = np.ascontiguousarray(np.asarray(vectors, dtype=np.float32))
        if tuple(emb.shape) != (n, self.dim):
            raise DenseEngineError(f"embeddings shape {tuple(emb.shape)} != ({n}, {self.dim})")
        ids_np = np.ascontiguousarray(np.asarray(ids, dtype=np.int64))
        if np.unique(ids_np).size != n:
            raise DenseEngineError("update_embeddings: ids must be distinct")
        _ffi.check(_ffi.lib().cdr_store_update_embeddings(self.handle, _ffi.ptr(ids_np), _ffi.ptr(emb), n,
                                                          self._stream()), "cdr_store_update_embeddings")

    def pending_ids(self, limit: Optional[int] = None, call_id: Any = None) -> np.ndarray:
        """Ids of rows whose embedding IS NULL, ascending -- the row set of _fetch_pending_rows
        (app/embedding_pipeline.py:121-146; the text-not-empty test is the caller's, texts live in the
        payload), optionally restricted to one call."""
        rows = self.rows
        valid = np.empty(rows, dtype=np.uint8)
        if rows:
            _ffi.check(_ffi.lib().cdr_store_read_valid(self.handle, 0, rows, _ffi.ptr(valid)), "cdr_store_read_valid")
        cols = self.host_columns()
        keep = valid == 0
        if call_id is not None:
            slot = self.slot_of_call(call_id)
            if slot is None and self.synthetic is not None and isinstance(call_id, (int, np.integer)):
                slot = int(call_id)
            keep &= cols["call_slot"] == (-1 if slot is None else slot)
        out = cols["ids"][keep]
        return out if limit is None else out[:limit]

    def append_literals(self, literals: Sequence[Optional[str]], ids: Sequence[int], **columns) -> None:
        """Append rows given as pgvector text literals -- the wire format the reference writes with
        ``UPDATE ... SET embedding = CAST(:embedding AS vector(D))`` (app/embedding_pipeline.py:149-168,
        literal built by _vector_literal, :63-64).  ``None`` models ``embedding IS NULL``.  Each element
        is parsed with float32 rounding, as pgvector's vector_in (strtof) does."""
        n = len(ids)
        emb = np.zeros((n, self.dim), dtype=np.float32)
        valid = np.ones(n, dtype=bool)
        for i, lit in enumerate(literals):
            if lit is None:
                valid[i] = False
                continue
            body = lit.strip()
            if not (body.startswith("[") and body.endswith("]")):
                raise DenseEngineError(f"row {i}: malformed vector literal")
            vals = np.array(body[1:-1].split(","), dtype=np.float32)
            if vals.shape[0] != self.dim:
                raise DenseEngineError(f"row {i}: expected {self.dim} dimensions, not {vals.shape[0]}")
            emb[i] = vals
        if "valid" in columns:
            valid &= np.asarray(columns.pop("valid"), dtype=bool)
        self.append(emb, ids, valid=valid, **columns)

    def append_embed_response(self, body: Dict[str, Any], ids: Sequence[int], **columns) -> None:
        """Append the ``{"embeddings": [[...], ...]}`` payload of the gateway's POST /embed
        (app/embeddings.py:71-82), validated like the reference client does."""
        vecs = body.get("embeddings")
        if not isinstance(vecs, list) or len(vecs) != len(ids):
            raise DenseEngineError("embedding response count mismatch")
        for i, v in enumerate(vecs):
            if len(v) != self.dim:
                raise DenseEngineError(f"embedding {i} has dim {len(v)}; expected {self.dim}")
        self.append(np.asarray(vecs, dtype=np.float32), ids, **columns)

    # ------------------------------------------------------------------ snapshot / restore
    def save(self, path: str, chunk_rows: int = 1 << 16) -> None:
        """Snapshot the resident store to a directory: raw fp32 rows (or bf16 bits for bf16-only
        stores) plus the filter columns, validity, dictionaries and payload.  The reference keeps this
        state durable in Postgres; here it is what a restart reloads instead of re-reading the DB."""
        import json, os
        os.makedirs(path, exist_ok=True)
        rows = self.rows
        what = "f32" if self.has_fp32 else "bf16"
        dt = np.float32 if self.has_fp32 else np.uint16
        mm = np.lib.format.open_memmap(os.path.join(path, f"emb_{what}.npy"), mode="w+", dtype=dt, shape=(rows, self.dim))
        valid = np.empty(rows, dtype=np.uint8)
        for r0 in range(0, rows, chunk_rows):
            m = min(chunk_rows, rows - r0)
            mm[r0:r0 + m] = self.read_rows(r0, m, (what,))[what]
            vslice = valid[r0:r0 + m]
            _ffi.check(_ffi.lib().cdr_store_read_valid(self.handle, r0, m, _ffi.ptr(vslice)), "cdr_store_read_valid")
        mm.flush(); del mm
        cols = self.read_rows(0, rows, ("ids", "call_slot", "started_at", "tag_bits"))
        np.savez(os.path.join(path, "columns.npz"), valid=valid, **cols)
        meta = dict(table_name=self.table_name, key_field=self.key_field, dim=self.dim, rows=rows, fp32=self.has_fp32,
                    bf16=self.has_bf16, tag_bits=self.tag_bits, synthetic=self.synthetic,
                    overflow_tag_slots={t: sorted(v) for t, v in self.overflow_tag_slots.items()},
                    slot_tag_mask={str(k): v for k, v in self.slot_tag_mask.items()})
        with open(os.path.join(path, "meta.json"), "w") as f:
            json.dump(meta, f)
        # host dictionaries as JSON (a snapshot directory is data, never code: nothing in it is unpickled).  Call ids
        # keep their type through a tag: UUIDs, ints and strings are what the reference's schema can hold.
        with open(os.path.join(path, "host_state.json"), "w") as f:
            json.dump(dict(call_ids_by_slot=[_encode_call_id(c) for c in self.call_ids_by_slot],
                           payload={str(k): v for k, v in self.payload.items()}), f, default=str)

    @classmethod
    def load(cls, path: str, device: Optional[int] = None, capacity_rows: Optional[int] = None,
             chunk_rows: int = 1 << 16) -> "DenseStore":
        """Rebuild a finalized store from :meth:`save`'s directory.  bf16-only snapshots restore the
        bf16 values exactly (they are widened to fp32, whose re-normalisation and RN-even rounding
        reproduce the same bits for already-rounded unit rows only approximately, so fp32 snapshots
        are the lossless form)."""
        import json, os
        torch = _torch()
        with open(os.path.join(path, "meta.json")) as f:
            meta = json.load(f)
        rows = meta["rows"]
        store = cls(meta["table_name"], capacity_rows or max(rows, 1), dim=meta["dim"], device=device,
                    fp32=meta["fp32"], bf16=meta["bf16"], key_field=meta["key_field"])
        cols = np.load(os.path.join(path, "columns.npz"))
        what = "f32" if meta["fp32"] else "bf16"
        mm = np.load(os.path.join(path, f"emb_{what}.npy"), mmap_mode="r")
        with open(os.path.join(path, "host_state.json")) as f:
            host = json.load(f)
        store.call_ids_by_slot = [_decode_call_id(c) for c in host["call_ids_by_slot"]]
        store.call_slots = {c: i for i, c in enumerate(store.call_ids_by_slot)}
        store.payload = {int(k): v for k, v in host["payload"].items()}
        store.tag_bits = {k: int(v) for k, v in meta["tag_bits"].items()}
        store.overflow_tag_slots = {t: set(v) for t, v in meta.get("overflow_tag_slots", {}).items()}
        store.slot_tag_mask = {int(k): int(v) for k, v in meta.get("slot_tag_mask", {}).items()}
        store.synthetic = meta["synthetic"]
        with torch.cuda.device(store.device):
            for r0 in range(0, rows, chunk_rows):
                m = min(chunk_rows, rows - r0)
                block = np.ascontiguousarray(mm[r0:r0 + m])
                if what == "bf16":
                    block = (block.astype(np.uint32) << 16).view(np.float32)
                sl = slice(r0, r0 + m)
                # keep the contiguous copies alive across the C call (ptr() only returns an address)
                c_ids = np.ascontiguousarray(cols["ids"][sl])
                c_slot = np.ascontiguousarray(cols["call_slot"][sl])
                c_time = np.ascontiguousarray(cols["started_at"][sl])
                c_tags = np.ascontiguousarray(cols["tag_bits"][sl])
                c_valid = np.ascontiguousarray(cols["valid"][sl])
                _ffi.check(_ffi.lib().cdr_store_append(
                    store.handle, _ffi.ptr(block), _ffi.ptr(c_ids), _ffi.ptr(c_slot), _ffi.ptr(c_time),
                    _ffi.ptr(c_tags), _ffi.ptr(c_valid), m, 0, store._stream()), "cdr_store_append")
        store.finalize()
        return store

    def append_synthetic(self, n: int, *, seed: int = SYNTH_CORPUS_SEED, first_row: int = 0,
                         id_base: int = 1, rows_per_call: int = SYNTH_ROWS_PER_CALL,
                         t0_us: int = SYNTH_T0_US, call_period_us: int = SYNTH_CALL_PERIOD_US) -> None:
        """Fill with rows first_row..first_row+n-1 of the synthetic corpus, generated on device
        (spec: oracle/synth_ref.c; BASELINE.md "Data")."""
        torch = _torch()
        with torch.cuda.device(self.device):
            _ffi.check(_ffi.lib().cdr_store_append_synthetic(
                self.handle, ctypes.c_uint64(seed), first_row, n, id_base, rows_per_call, t0_us,
                call_period_us, self._stream()), "cdr_store_append_synthetic")
        self.synthetic = dict(seed=seed, first_row=first_row, id_base=id_base, rows_per_call=rows_per_call,
                              t0_us=t0_us, call_period_us=call_period_us)
        # synthetic call ids are the slot numbers themselves; 16 synthetic tags "tag0".."tag15"
        if not self.tag_bits:
            self.tag_bits = {f"tag{i}": i for i in range(16)}

    def finalize(self) -> None:
        torch = _torch()
        with torch.cuda.device(self.device):
            _ffi.check(_ffi.lib().cdr_store_finalize(self.handle, self._stream()), "cdr_store_finalize")
        self.finalized = True

    # ------------------------------------------------------------------ introspection
    def info(self) -> Dict[str, int]:
        rows, dim, flags, nv, dev = (ctypes.c_int64(), ctypes.c_int32(), ctypes.c_uint32(),
                                     ctypes.c_int64(), ctypes.c_int32())
        _ffi.check(_ffi.lib().cdr_store_info(self.handle, ctypes.byref(rows), ctypes.byref(dim),
                                             ctypes.byref(flags), ctypes.byref(nv), ctypes.byref(dev)))
        return dict(rows=rows.value, dim=dim.value, flags=flags.value, n_valid=nv.value, device=dev.value)

    @property
    def rows(self) -> int:
        if self.finalized:                      # the store is immutable after finalize
            if self._rows is None:
                self._rows = self.info()["rows"]
            return self._rows
        return self.info()["rows"]

    @property
    def has_fp32(self) -> bool:
        return bool(self.flags & _ffi.CDR_STORE_FP32)

    @property
    def has_bf16(self) -> bool:
        return bool(self.flags & _ffi.CDR_STORE_BF16)

    def read_rows(self, first_row: int, n: int, what: Sequence[str] = ("f32",)) -> Dict[str, np.ndarray]:
        """Copy resident columns back to the host (tests / snapshots)."""
        out: Dict[str, np.ndarray] = {}
        spec = {"f32": ((n, self.dim), np.float32), "bf16": ((n, self.dim), np.uint16), "ids": ((n,), np.int64),
                "call_slot": ((n,), np.int32), "started_at": ((n,), np.int64), "tag_bits": ((n,), np.uint64),
                "inv_norm": ((n,), np.float32)}
        order = ["f32", "bf16", "ids", "call_slot", "started_at", "tag_bits", "inv_norm"]
        ptrs = []
        for name in order:
            if name in what:
                shape, dt = spec[name]
                out[name] = np.empty(shape, dtype=dt)
                ptrs.append(_ffi.ptr(out[name]))
            else:
                ptrs.append(None)
        _ffi.check(_ffi.lib().cdr_store_read_rows(self.handle, first_row, n, *ptrs), "cdr_store_read_rows")
        return out

    def read_rows_device(self, first_row: int, n: int, what: str = "f32"):
        """Rows [first_row, first_row+n) as stored, copied into a new CUDA tensor on the current stream
        (``cdr_store_copy_rows_device``): "f32" -> float32 [n, dim], "bf16" -> bfloat16 [n, dim]."""
        torch = _torch()
        if what not in ("f32", "bf16"):
            raise ValueError(f"read_rows_device: {what!r}: expected 'f32' or 'bf16'")
        out = torch.empty((n, self.dim), dtype=torch.float32 if what == "f32" else torch.bfloat16,
                          device=f"cuda:{self.device}")
        _ffi.check(_ffi.lib().cdr_store_copy_rows_device(
            self.handle, first_row, n, _ffi.ptr(out) if what == "f32" else None,
            _ffi.ptr(out) if what == "bf16" else None, self._stream()), "cdr_store_copy_rows_device")
        return out

    def host_columns(self) -> Dict[str, np.ndarray]:
        """ids / call_slot / started_at / tag_bits of all rows on the host (cached; lexical lane)."""
        if self._host_cols is None:
            self._host_cols = self.read_rows(0, self.rows, ("ids", "call_slot", "started_at", "tag_bits"))
        return self._host_cols

    # ------------------------------------------------------------------ filters (K6)
    def slot_bitmap(self, call_slots: Optional[Sequence[int]]):
        """Host bitmap over call slots for ``call_id = ANY(:call_ids)``: (uint32 words or None, n_slots).
        None = unscoped; an all-zero bitmap = ``call_ids == []``."""
        if call_slots is None:
            return None, 0
        n_slots = max(len(self.call_ids_by_slot), (max(call_slots) + 1) if len(call_slots) else 0, 1)
        if self.synthetic is not None:
            n_slots = max(n_slots, (self.synthetic["first_row"] + self.rows) // self.synthetic["rows_per_call"] + 1)
        bm = np.zeros((n_slots + 31) // 32, dtype=np.uint32)
        for s in call_slots:
            if 0 <= s < n_slots:
                bm[s >> 5] |= np.uint32(1 << (s & 31))
        return bm, n_slots

    def filter_spec_struct(self, *, call_slots=None, date_from=None, date_to=None, tag_mask=None, dense_lane: int = 0):
        """(cdr_filter_spec, keep-alive) for the fused C call; None when the filter is empty and the dense lane
        is the default exact scan.  dense_lane: _ffi.CDR_DENSE_LANE_* (the batched bf16 lane for mode "ann")."""
        if call_slots is None and date_from is None and date_to is None and tag_mask is None and not dense_lane:
            return None, None
        bm, n_slots = self.slot_bitmap(call_slots)
        spec = _ffi.FilterSpec()
        spec.dense_lane = int(dense_lane)
        spec.call_slot_bitmap_host = _ffi.ptr(bm)
        spec.n_call_slots = n_slots
        spec.has_date_from = 0 if date_from is None else 1
        spec.date_from_us = 0 if date_from is None else to_micros(date_from)
        spec.has_date_to = 0 if date_to is None else 1
        spec.date_to_us = 0 if date_to is None else to_micros(date_to)
        spec.has_tag_filter = 0 if tag_mask is None else 1
        spec.tag_any = int(tag_mask or 0)
        return spec, bm

    def filter_bitmap(self, *, call_slots: Optional[Sequence[int]] = None, date_from=None, date_to=None,
                      tag_mask: Optional[int] = None):
        """Build the allow-bitmap for a WHERE clause and count its rows.

        call_slots: None = unscoped; [] = ``call_ids == []`` (matches nothing).
        tag_mask: None = no tag filter; 0 = a tag filter whose tags are all unknown (matches nothing).
        Returns (allow: torch.int32 CUDA tensor [ceil(capacity/32)], count: int)."""
        torch = _torch()
        # the bitmap covers the store's CAPACITY: rows appended after it was built read as "not allowed", and a scan
        # that already sees the larger row count never reads past it
        words = (self.capacity + 31) // 32
        bm, n_slots = self.slot_bitmap(call_slots)
        count = ctypes.c_int64(0)
        with torch.cuda.device(self.device):
            allow = torch.empty(max(words, 1), dtype=torch.int32, device=f"cuda:{self.device}")
            _ffi.check(_ffi.lib().cdr_filter_build(
                self.handle, _ffi.ptr(bm), n_slots,
                0 if date_from is None else 1, 0 if date_from is None else to_micros(date_from),
                0 if date_to is None else 1, 0 if date_to is None else to_micros(date_to),
                0 if tag_mask is None else 1, ctypes.c_uint64(tag_mask or 0),
                _ffi.ptr(allow), ctypes.byref(count), self._stream()), "cdr_filter_build")
        return allow, int(count.value)

    # ------------------------------------------------------------------ search
    def _search(self, fn_name: str, queries, k: int, allow) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        torch = _torch()
        on_dev = hasattr(queries, "is_cuda") and queries.is_cuda
        if on_dev:
            q = queries.to(dtype=torch.float32).contiguous()
            if q.dim() == 1:
                q = q.unsqueeze(0)
        else:
            q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
            if q.ndim == 1:
                q = q[None, :]
        nq = int(q.shape[0])
        if int(q.shape[1]) != self.dim:
            raise DenseEngineError(f"query dim {int(q.shape[1])} != store dim {self.dim}")
        # no torch.cuda.device() context here: every tensor names its device, the stream is the store
        # device's current stream, and the C entry points switch devices themselves (DeviceGuard)
        if on_dev:
            dev = q.device
            if dev.index != self.device:
                raise DenseEngineError(f"queries live on {dev}, the store on cuda:{self.device}")
            # One allocation for the three outputs, addressed by pointer arithmetic: the launches are enqueued before any
            # view is built (a single-query request is a few hundred microseconds of kernels, and every host operation
            # ahead of the first launch is GPU idle time in its latency); the views are made while the GPU works.
            buf = torch.empty((2 * nq * k + (nq + 1) // 2,), dtype=torch.int64, device=dev)
            base = buf.data_ptr()
            fn = getattr(_ffi.lib(), fn_name)
            _ffi.check(fn(self.handle, _ffi.ptr(q), nq, k, _ffi.ptr(allow), base, base + nq * k * 8, base + 2 * nq * k * 8,
                          self._stream()), fn_name)
            sc = buf[:nq * k].view(torch.float64).view(nq, k)
            ids = buf[nq * k:2 * nq * k].view(nq, k)
            cnt = buf[2 * nq * k:].view(torch.int32)[:nq]
            return ids, sc, cnt
        sc = np.empty((nq, k), dtype=np.float64)
        ids = np.empty((nq, k), dtype=np.int64)
        cnt = np.empty((nq,), dtype=np.int32)
        fn = getattr(_ffi.lib(), fn_name + "_host")
        _ffi.check(fn(self.handle, _ffi.ptr(q), nq, k, _ffi.ptr(allow), _ffi.ptr(sc), _ffi.ptr(ids),
                      _ffi.ptr(cnt), self._stream()), fn_name + "_host")
        return ids, sc, cnt

    def search_exact(self, queries, k: int, allow=None, shared: bool = False):
        """mode="exact": fp32 cosine scan (K1) + fp64 re-score.  Host queries -> host results
        (numpy, through the *_host C entry point: H2D + kernels + D2H + sync); CUDA tensors ->
        CUDA tensors (asynchronous on the current stream).  Returns (ids[nq,k], scores[nq,k], n[nq]).
        shared=True: a batch of concurrent requests shares every tile read among 3 queries (16 for larger batches)
        (``cdr_search_exact_f32_shared``: nq/3 scans of the corpus, same bits); False = one scan per query."""
        return self._search("cdr_search_exact_f32_shared" if shared else "cdr_search_exact_f32", queries, k, allow)

    def search_scan_bf16(self, queries, k: int, allow=None):
        """mode="ann" for a single query or a few (``cdr_search_scan_bf16``): the HBM-bound scan over the bf16 copy
        of the rows (half the bytes of the exact scan), candidate lists twice as wide, exact re-score."""
        return self._search("cdr_search_scan_bf16", queries, k, allow)

    def search_batch(self, queries, k: int, allow=None):
        """mode="ann" served by the batched bf16 tensor-core lane (K2) + exact re-score."""
        return self._search("cdr_search_batch_bf16", queries, k, allow)

    # ------------------------------------------------------------------ fused hybrid /retrieve
    def hybrid_retrieve(self, queries, dense_k: int, *, tech_index=None, token_ids=None, n_tokens=None,
                        tech_limit: int = 50, bm25_ids=None, bm25_offsets=None, rrf_k: int = 60,
                        filter_spec: Optional[Dict[str, Any]] = None, max_out: Optional[int] = None,
                        filter_specs: Optional[Sequence[Optional[Dict[str, Any]]]] = None,
                        group_offsets: Optional[Sequence[int]] = None):
        """One ``cdr_hybrid_retrieve_host`` call (include/cadence_dense.h): filter -> dense exact lane ->
        tech_tokens lane -> RRF for nq queries sharing one filter; one H2D, one D2H, one sync.

        queries: [nq, dim] float32 numpy or None (dense lane disabled).  token_ids [nq, T] int32 /
        n_tokens [nq] int32 with ``tech_index`` a DeviceTechIndex.  bm25_ids / bm25_offsets: the opaque
        BM25 lane (ranked ids back to back, offsets [nq+1]).  Returns a dict of numpy arrays.

        Requests with different filters: pass ``filter_specs`` (one spec per group, None = unscoped) and
        ``group_offsets`` ([n_groups + 1], queries of a group are consecutive) instead of ``filter_spec``
        (``cdr_hybrid_retrieve_groups_host``); ``count`` is then a list with one COUNT(*) per group."""
        torch = _torch()
        dense = queries is not None
        if dense:
            q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
            if q.ndim == 1:
                q = q[None, :]
            if q.shape[1] != self.dim:
                raise DenseEngineError(f"query dim {q.shape[1]} != store dim {self.dim}")
            nq = int(q.shape[0])
        else:
            q = None
            nq = int(len(n_tokens)) if n_tokens is not None else (len(bm25_offsets) - 1 if bm25_offsets is not None else 0)
        tok = nt = None
        max_tokens = 1
        if tech_index is not None and token_ids is not None:
            tok = np.ascontiguousarray(token_ids, dtype=np.int32)
            nt = np.ascontiguousarray(n_tokens, dtype=np.int32)
            max_tokens = int(tok.shape[1])
        b_ids = b_off = None
        bm25_max = 0
        if bm25_offsets is not None:
            b_off = np.ascontiguousarray(bm25_offsets, dtype=np.int32)
            b_ids = np.ascontiguousarray(bm25_ids if bm25_ids is not None else np.empty(0), dtype=np.int64)
            bm25_max = int(np.diff(b_off).max()) if nq else 0
        kd = int(dense_k) if dense else 0
        if max_out is None:
            max_out = max(1, bm25_max + int(tech_limit) + kd)
        grouped = filter_specs is not None
        if grouped:
            n_groups = len(filter_specs)
            offs = np.ascontiguousarray(group_offsets, dtype=np.int32)
            if offs.shape[0] != n_groups + 1:
                raise DenseEngineError("group_offsets must have one more entry than filter_specs")
            specs = (_ffi.FilterSpec * n_groups)()
            keep = []
            for gi, fs in enumerate(filter_specs):
                one, alive = self.filter_spec_struct(**(fs or {}))
                keep.append(alive)
                if one is not None:
                    specs[gi] = one
        else:
            spec, keep = self.filter_spec_struct(**(filter_spec or {}))
        out = {"dense_ids": np.empty((nq, max(kd, 1)), dtype=np.int64), "dense_scores": np.empty((nq, max(kd, 1)), dtype=np.float64),
               "dense_n": np.zeros(nq, dtype=np.int32), "tech_ids": np.empty((nq, tech_limit), dtype=np.int64),
               "tech_n": np.zeros(nq, dtype=np.int32), "fused_ids": np.empty((nq, max_out), dtype=np.int64),
               "fused_scores": np.empty((nq, max_out), dtype=np.float64), "fused_mask": np.empty((nq, max_out), dtype=np.uint32),
               "fused_n": np.zeros(nq, dtype=np.int32)}
        if grouped:
            counts = np.zeros(n_groups, dtype=np.int64)
            _ffi.check(_ffi.lib().cdr_hybrid_retrieve_groups_host(
                self.handle, None if tech_index is None else tech_index._h, ctypes.addressof(specs), _ffi.ptr(offs), n_groups,
                _ffi.ptr(q), nq, kd, _ffi.ptr(tok), _ffi.ptr(nt), max_tokens, int(tech_limit), _ffi.ptr(b_ids), _ffi.ptr(b_off),
                int(rrf_k), int(max_out), _ffi.ptr(counts), _ffi.ptr(out["dense_ids"]), _ffi.ptr(out["dense_scores"]),
                _ffi.ptr(out["dense_n"]), _ffi.ptr(out["tech_ids"]), _ffi.ptr(out["tech_n"]), _ffi.ptr(out["fused_ids"]),
                _ffi.ptr(out["fused_scores"]), _ffi.ptr(out["fused_mask"]), _ffi.ptr(out["fused_n"]), self._stream()),
                "cdr_hybrid_retrieve_groups_host")
            out["count"] = counts.tolist()
            return out
        count = ctypes.c_int64(0)
        _ffi.check(_ffi.lib().cdr_hybrid_retrieve_host(
            self.handle, None if tech_index is None else tech_index._h, None if spec is None else ctypes.addressof(spec),
            _ffi.ptr(q), nq, kd, _ffi.ptr(tok), _ffi.ptr(nt), max_tokens, int(tech_limit), _ffi.ptr(b_ids), _ffi.ptr(b_off),
            int(rrf_k), int(max_out), ctypes.addressof(count), _ffi.ptr(out["dense_ids"]), _ffi.ptr(out["dense_scores"]),
            _ffi.ptr(out["dense_n"]), _ffi.ptr(out["tech_ids"]), _ffi.ptr(out["tech_n"]), _ffi.ptr(out["fused_ids"]),
            _ffi.ptr(out["fused_scores"]), _ffi.ptr(out["fused_mask"]), _ffi.ptr(out["fused_n"]), self._stream()),
            "cdr_hybrid_retrieve_host")
        out["count"] = int(count.value)
        return out
