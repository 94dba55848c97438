"""Embedding client with the call surface of the reference's app/embeddings.py.

  embeddings_enabled()            app/embeddings.py:21-22
  embed_texts(texts)              app/embeddings.py:48-82   POST {base}/embed {"texts", "model"}
  embed_texts_batched(texts, n)   app/embeddings.py:85-100
  EmbeddingResult / EmbeddingClientError   app/embeddings.py:11-18

The remote model server is out of scope.  What is kept is the
contract: which texts are sent, how a reply is validated, which message each failure carries.
Two back ends produce vectors: the HTTP gateway (`_Gateway`, used when EMBEDDINGS_BASE_URL is
set) and an in-process callable installed with :func:`set_embedder` (benches and tests use
:class:`SyntheticEmbedder`); both go through the same validation.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Iterator, List, Optional, Sequence, Tuple

from .config import settings

try:  # what the reference talks HTTP with; only needed when a base URL is configured
    import httpx
except Exception:  # pragma: no cover
    httpx = None  # type: ignore

_DETAIL_LIMIT = 400      # characters of an error body quoted back to the caller


class EmbeddingClientError(RuntimeError):
    """Any failure of the embedding step; retrieval catches it and serves lexical-only."""


@dataclass(frozen=True)
class EmbeddingResult:
    vectors: List[List[float]]
    model: str


Embedder = Callable[[List[str]], "EmbeddingResult"]
_embedder: Optional[Embedder] = None


def set_embedder(fn: Optional[Embedder]) -> None:
    """Install (or, with None, remove) an in-process embedder that replaces the HTTP gateway."""
    global _embedder
    _embedder = fn


def embeddings_enabled() -> bool:
    if _embedder is not None:
        return True
    return settings.embeddings_base_url.strip() != ""


# ------------------------------------------------------------------ validation shared by both back ends
def _validate_texts(texts: Sequence[str]) -> List[str]:
    kept = []
    for item in texts:
        if isinstance(item, str):
            stripped = item.strip()
            if stripped:
                kept.append(stripped)
    if len(kept) == 0:
        raise EmbeddingClientError("embedding request requires at least one non-empty text")
    return kept


def _validate_vectors(vectors: Sequence[Sequence[float]]) -> List[List[float]]:
    dim = settings.embeddings_dim
    for position, row in enumerate(vectors):
        if len(row) != dim:
            raise EmbeddingClientError(f"embedding {position} has dim {len(row)}; expected {dim}")
    return [list(map(float, row)) for row in vectors]


def _checked(vectors: Any, n_texts: int, model: Optional[str]) -> EmbeddingResult:
    if len(vectors) != n_texts:
        raise EmbeddingClientError(f"embedding response count mismatch: got {len(vectors)}, expected {n_texts}")
    return EmbeddingResult(vectors=_validate_vectors(vectors), model=str(model or settings.embeddings_model_id))


# ------------------------------------------------------------------ HTTP gateway
class _Gateway:
    """`POST {base}/embed` with body {"texts": [...], "model": id}; reply {"embeddings": [[...]], "model": id}."""

    def __init__(self) -> None:
        self.url = settings.embeddings_base_url.rstrip("/") + "/embed"
        self.model_id = settings.embeddings_model_id
        self.timeout_s = settings.embeddings_timeout_s

    def __call__(self, texts: List[str]) -> Tuple[Any, Optional[str]]:
        if httpx is None:
            raise EmbeddingClientError("embedding HTTP request failed: httpx is not installed")
        try:
            with httpx.Client(timeout=httpx.Timeout(self.timeout_s)) as client:
                reply = client.post(self.url, json={"texts": texts, "model": self.model_id})
        except httpx.HTTPError as exc:
            raise EmbeddingClientError(f"embedding HTTP request failed: {exc}") from exc
        if reply.status_code != 200:
            raise EmbeddingClientError(
                f"embedding service returned {reply.status_code}: {reply.text.strip()[:_DETAIL_LIMIT]}")
        body = reply.json()
        vectors = body.get("embeddings")
        if not isinstance(vectors, list):
            raise EmbeddingClientError("embedding response missing 'embeddings' list")
        return vectors, body.get("model")


def embed_texts(texts: Sequence[str]) -> EmbeddingResult:
    if not embeddings_enabled():
        raise EmbeddingClientError("EMBEDDINGS_BASE_URL is not configured")
    wanted = _validate_texts(texts)
    if _embedder is not None:
        produced = _embedder(wanted)
        return _checked(produced.vectors, len(wanted), produced.model)
    vectors, model = _Gateway()(wanted)
    return _checked(vectors, len(wanted), model)


def _windows(items: List[str], size: int) -> Iterator[List[str]]:
    for first in range(0, len(items), size):
        yield items[first:first + size]


def embed_texts_batched(texts: Sequence[str], batch_size: Optional[int] = None) -> EmbeddingResult:
    wanted = _validate_texts(texts)
    size = batch_size or settings.embeddings_batch_size
    if size <= 0:
        raise EmbeddingClientError("batch size must be > 0")
    rows: List[List[float]] = []
    model = settings.embeddings_model_id
    for window in _windows(wanted, size):
        part = embed_texts(window)
        rows += part.vectors
        model = part.model          # the last batch names the model, as in the reference
    return EmbeddingResult(vectors=rows, model=model)


# ------------------------------------------------------------------ in-process stand-in for the gateway
class SyntheticEmbedder:
    """Deterministic stand-in for the gateway: text -> row of the synthetic query stream (same
    generator as the corpus, seed = query seed).  The row index is a stable 63-bit
    hash of the text, so equal texts embed equally.  Vectors are produced on the GPU by the engine's
    generator and handed back as Python floats, like the HTTP client does."""

    def __init__(self, seed: int = 20260210, dim: Optional[int] = None, model: str = "synthetic-philox"):
        self.seed = seed
        self.dim = dim or settings.embeddings_dim
        self.model = model

    @staticmethod
    def _row_of(text: str) -> int:
        import hashlib
        return int.from_bytes(hashlib.sha256(text.encode("utf-8")).digest()[:8], "little") >> 1

    def __call__(self, texts: List[str]) -> EmbeddingResult:
        from .store import synth_rows_device
        vecs = [synth_rows_device(self.seed, self._row_of(t), 1, self.dim).cpu().tolist()[0] for t in texts]
        return EmbeddingResult(vectors=vecs, model=self.model)
