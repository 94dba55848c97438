"""Write side of the dense lane over the resident store -- the reference's app/embedding_pipeline.py
(`embed_backfill`) with the SQL replaced by store calls:

  _fetch_pending_rows   :121-146   rows WHERE embedding IS NULL AND text not empty, ORDER BY id LIMIT n
  _embed_texts_adaptive :83-118    embed in batches, halve the batch when the provider rejects its size
  _update_embeddings    :149-168   UPDATE ... SET embedding = CAST(:e AS vector(D)) WHERE id = :row_id
  _backfill_table       :210-238   loop until no row is pending
  run_embedding_backfill:241-282   all tables -> BackfillSummary

The texts come from the payload registered with the store (`text` for chunks, `content` for
artifact_chunks -- the reference's TableSpec.text_column).  The ingestion_runs bookkeeping
(:171-207) is Postgres-side metadata and stays there."""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Set, Tuple

from .config import settings
from .embeddings import EmbeddingClientError, EmbeddingResult, embed_texts, embeddings_enabled
from .store import DenseStore


@dataclass
class PendingRow:
    row_id: int
    call_id: Any
    content: str


@dataclass
class BackfillSummary:
    rows_updated: int
    calls_touched: int
    model_used: str
    per_table: Dict[str, int] = field(default_factory=dict)


_TEXT_COLUMN = {"chunks": "text", "artifact_chunks": "content"}


# provider messages that name their batch limit: "... batch-size must be <= 8 ...", "maximum batch size ... 16"
_BATCH_SIZE_LIMIT_PATTERNS = (
    re.compile(r"batch[- ]size[^0-9]{0,40}<=\s*(\d+)", re.IGNORECASE),
    re.compile(r"max(?:imum)?\s+batch[- ]size[^0-9]{0,40}(\d+)", re.IGNORECASE),
)


def infer_batch_size_limit(error_message: str) -> Optional[int]:
    """Provider error text -> the batch limit it names, if any (app/embedding_pipeline.py:57-85)."""
    message = (error_message or "").strip()
    if not message:
        return None
    for pattern in _BATCH_SIZE_LIMIT_PATTERNS:
        m = pattern.search(message)
        if not m:
            continue
        try:
            value = int(m.group(1))
        except (TypeError, ValueError):
            continue
        if value > 0:
            return value
    return None


def _embed_texts_adaptive(texts: Sequence[str], batch_size: int) -> EmbeddingResult:
    cleaned = list(texts)
    vectors: List[List[float]] = []
    model_used = settings.embeddings_model_id
    current_batch = max(1, batch_size)
    index = 0
    while index < len(cleaned):
        upper = min(len(cleaned), index + current_batch)
        chunk = cleaned[index:upper]
        try:
            result = embed_texts(chunk)
        except EmbeddingClientError as exc:
            if len(chunk) <= 1:
                raise
            inferred = infer_batch_size_limit(str(exc))
            current_batch = max(1, inferred) if inferred is not None and inferred < len(chunk) else max(1, len(chunk) // 2)
            continue
        vectors.extend(result.vectors)
        model_used = result.model
        index = upper
    return EmbeddingResult(vectors=vectors, model=model_used)


def _fetch_pending_rows(store: DenseStore, limit: int, call_id: Any = None) -> List[PendingRow]:
    text_column = _TEXT_COLUMN.get(store.table_name, "text")
    cols = store.host_columns()
    out: List[PendingRow] = []
    for row_id in store.pending_ids(None, call_id).tolist():
        content = store.payload.get(row_id, {}).get(text_column)
        if content is None or not str(content).strip():
            continue                                   # AND text IS NOT NULL AND length(trim(text)) > 0
        pos = int(cols["ids"].searchsorted(row_id))
        slot = int(cols["call_slot"][pos])
        call = store.call_ids_by_slot[slot] if slot < len(store.call_ids_by_slot) else slot
        out.append(PendingRow(row_id=row_id, call_id=call, content=str(content)))
        if len(out) >= limit:
            break
    return out


def _update_embeddings(store: DenseStore, rows: Sequence[PendingRow], vectors: Sequence[Sequence[float]]) -> None:
    if len(rows) != len(vectors):
        raise RuntimeError(f"row/vector mismatch for {store.table_name}: {len(rows)} rows vs {len(vectors)} vectors")
    store.update_embeddings([r.row_id for r in rows], vectors)


def _backfill_table(store: DenseStore, *, batch_size: int, call_id: Any = None) -> Tuple[int, Set[Any], str]:
    updated = 0
    touched: Set[Any] = set()
    model_used = settings.embeddings_model_id
    while True:
        batch = _fetch_pending_rows(store, batch_size, call_id=call_id)
        if not batch:
            break
        result = _embed_texts_adaptive([row.content for row in batch], batch_size=batch_size)
        _update_embeddings(store, batch, result.vectors)
        touched.update(row.call_id for row in batch)
        updated += len(batch)
        model_used = result.model
    return updated, touched, model_used


def run_embedding_backfill(stores: Sequence[DenseStore], *, batch_size: int, call_id: Any = None) -> BackfillSummary:
    if not embeddings_enabled():
        raise RuntimeError("EMBEDDINGS_BASE_URL must be set to run embedding backfill")
    if settings.embeddings_dim <= 0:
        raise RuntimeError("EMBEDDINGS_DIM must be > 0")
    if batch_size <= 0:
        raise RuntimeError("EMBEDDINGS_BATCH_SIZE must be > 0")
    total = 0
    calls: Set[Any] = set()
    model_used = settings.embeddings_model_id
    per_table: Dict[str, int] = {}
    for store in stores:
        updated, touched, model_used = _backfill_table(store, batch_size=batch_size, call_id=call_id)
        per_table[store.table_name] = updated
        total += updated
        calls |= touched
    return BackfillSummary(rows_updated=total, calls_touched=len(calls), model_used=model_used, per_table=per_table)
