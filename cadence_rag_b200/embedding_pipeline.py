"""Write side of the dense lane over the resident store.

The reference backfills embeddings with `embed_backfill` (app/embedding_pipeline.py): rows are ingested
with `embedding IS NULL`, later selected in id order, embedded in batches that shrink when the
provider rejects their size, and written back one `UPDATE ... SET embedding = CAST(:e AS vector(D))`
per row.  Here the table is a `DenseStore`: the pending set comes from its validity bitmap and
payload texts, the write is `DenseStore.update_embeddings` (one kernel per batch, in place).

Same entry points as the reference so the call sites read alike:

  infer_batch_size_limit   :57-85     provider error text -> the batch limit it names
  _embed_texts_adaptive    :88-118    embed, shrinking the batch on provider errors
  _fetch_pending_rows      :121-146   WHERE embedding IS NULL AND text not blank ORDER BY id LIMIT n
  _update_embeddings       :149-168   the per-row UPDATE
  _backfill_table          :210-238   until nothing is pending
  run_embedding_backfill   :241-282   every table -> BackfillSummary

The ingestion_runs bookkeeping (:171-207) is Postgres-side metadata and stays there.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Any, Dict, Iterator, List, Optional, Sequence, Set, Tuple

from .config import settings
from .embeddings import EmbeddingClientError, EmbeddingResult, embed_texts, embeddings_enabled
from .store import DenseStore

# which payload column holds the text that gets embedded (the reference's TableSpec.text_column)
_TEXT_COLUMN = {"chunks": "text", "artifact_chunks": "content"}

# provider messages that name their limit: "... batch-size must be <= 8 ...", "maximum batch size ... 16"
_LIMIT_IN_MESSAGE = [re.compile(rx, re.IGNORECASE) for rx in (
    r"batch[- ]size[^0-9]{0,40}<=\s*(\d+)",
    r"max(?:imum)?\s+batch[- ]size[^0-9]{0,40}(\d+)",
)]


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


def infer_batch_size_limit(error_message: str) -> Optional[int]:
    text = (error_message or "").strip()
    for rx in _LIMIT_IN_MESSAGE if text else ():
        found = rx.search(text)
        if found is None:
            continue
        try:
            limit = int(found.group(1))
        except (TypeError, ValueError):
            continue
        if limit > 0:
            return limit
    return None


def _embed_texts_adaptive(texts: Sequence[str], batch_size: int) -> EmbeddingResult:
    """Embed `texts` in order.  A provider error on a batch of more than one text shrinks the batch -- to the
    limit the message names when it names a smaller one, else to half -- and the same position is retried; the
    smaller size sticks for the rest of the run.  An error on a single text is final."""
    pending = list(texts)
    size = max(1, batch_size)
    done = 0
    rows: List[List[float]] = []
    model = settings.embeddings_model_id
    while done < len(pending):
        window = pending[done:done + size]
        try:
            part = embed_texts(window)
        except EmbeddingClientError as exc:
            if len(window) <= 1:
                raise
            named = infer_batch_size_limit(str(exc))
            size = max(1, named if (named is not None and named < len(window)) else len(window) // 2)
        else:
            rows += part.vectors
            model = part.model
            done += len(window)
    return EmbeddingResult(vectors=rows, model=model)


class _TableBackfill:
    """The pending-rows / embed / update loop over one resident table."""

    def __init__(self, store: DenseStore, batch_size: int, call_id: Any = None):
        self.store = store
        self.batch_size = batch_size
        self.call_id = call_id
        self.text_column = _TEXT_COLUMN.get(store.table_name, "text")

    def _call_of(self, row_id: int) -> Any:
        cols = self.store.host_columns()
        slot = int(cols["call_slot"][int(cols["ids"].searchsorted(row_id))])
        known = self.store.call_ids_by_slot
        return known[slot] if slot < len(known) else slot

    def pending(self, limit: int) -> List[PendingRow]:
        picked: List[PendingRow] = []
        for row_id in self.store.pending_ids(None, self.call_id).tolist():
            content = self.store.payload.get(row_id, {}).get(self.text_column)
            if content is None or str(content).strip() == "":
                continue                      # text IS NOT NULL AND length(trim(text)) > 0
            picked.append(PendingRow(row_id=row_id, call_id=self._call_of(row_id), content=str(content)))
            if len(picked) == limit:
                break
        return picked

    def apply(self, rows: Sequence[PendingRow], vectors: Sequence[Sequence[float]]) -> None:
        if len(rows) != len(vectors):
            raise RuntimeError(f"row/vector mismatch for {self.store.table_name}: {len(rows)} rows vs {len(vectors)} vectors")
        self.store.update_embeddings([row.row_id for row in rows], vectors)

    def batches(self) -> Iterator[Tuple[List[PendingRow], EmbeddingResult]]:
        while True:
            rows = self.pending(self.batch_size)
            if not rows:
                return
            embedded = _embed_texts_adaptive([row.content for row in rows], batch_size=self.batch_size)
            self.apply(rows, embedded.vectors)
            yield rows, embedded

    def run(self) -> Tuple[int, Set[Any], str]:
        updated, calls, model = 0, set(), settings.embeddings_model_id
        for rows, embedded in self.batches():
            updated += len(rows)
            calls.update(row.call_id for row in rows)
            model = embedded.model
        return updated, calls, model


def _fetch_pending_rows(store: DenseStore, limit: int, call_id: Any = None) -> List[PendingRow]:
    return _TableBackfill(store, limit, call_id).pending(limit)


def _update_embeddings(store: DenseStore, rows: Sequence[PendingRow], vectors: Sequence[Sequence[float]]) -> None:
    _TableBackfill(store, max(1, len(rows))).apply(rows, vectors)


def _backfill_table(store: DenseStore, *, batch_size: int, call_id: Any = None) -> Tuple[int, Set[Any], str]:
    return _TableBackfill(store, batch_size, call_id).run()


def run_embedding_backfill(stores: Sequence[DenseStore], *, batch_size: int, call_id: Any = None) -> BackfillSummary:
    if not embeddings_enabled():
        raise RuntimeError("EMBEDDINGS_BASE_URL must be set to run embedding backfill")
    if settings.embeddings_dim <= 0:
        raise RuntimeError("EMBEDDINGS_DIM must be > 0")
    if batch_size <= 0:
        raise RuntimeError("EMBEDDINGS_BATCH_SIZE must be > 0")
    summary = BackfillSummary(rows_updated=0, calls_touched=0, model_used=settings.embeddings_model_id)
    touched: Set[Any] = set()
    for store in stores:
        updated, calls, summary.model_used = _backfill_table(store, batch_size=batch_size, call_id=call_id)
        summary.per_table[store.table_name] = updated
        summary.rows_updated += updated
        touched |= calls
    summary.calls_touched = len(touched)
    return summary
