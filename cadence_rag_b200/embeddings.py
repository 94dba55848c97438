"""Embedding client -- same call surface as the reference's app/embeddings.py.

  embeddings_enabled()            app/embeddings.py:21-22
  embed_texts(texts)              app/embeddings.py:48-82   POST {base}/embed {"texts","model"}
  embed_texts_batched(texts, n)   app/embeddings.py:85-100
  EmbeddingResult / EmbeddingClientError   app/embeddings.py:11-18

The remote model server is out of scope: the HTTP path is kept
verbatim in behaviour, and benches/tests install an in-process embedder with
:func:`set_embedder` (e.g. :class:`SyntheticEmbedder`) instead of a base URL.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

from .config import settings

try:  # httpx is what the reference uses; only needed when a base URL is configured
    import httpx
except Exception:  # pragma: no cover
    httpx = None  # type: ignore


class EmbeddingClientError(RuntimeError):
    pass


@dataclass(frozen=True)
class EmbeddingResult:
    vectors: List[List[float]]
    model: str


_embedder: Optional[Callable[[List[str]], "EmbeddingResult"]] = None


def set_embedder(fn: Optional[Callable[[List[str]], "EmbeddingResult"]]) -> None:
    """Install (or clear) an in-process embedder used instead of the HTTP gateway."""
    global _embedder
    _embedder = fn


def embeddings_enabled() -> bool:
    return _embedder is not None or bool(settings.embeddings_base_url.strip())


def _normalize_base_url(raw: str) -> str:
    return raw.rstrip("/")


def _validate_texts(texts: Sequence[str]) -> List[str]:
    cleaned = [t.strip() for t in texts if isinstance(t, str) and t.strip()]
    if not cleaned:
        raise EmbeddingClientError("embedding request requires at least one non-empty text")
    return cleaned


def _validate_vectors(vectors: Sequence[Sequence[float]]) -> List[List[float]]:
    want = settings.embeddings_dim
    out: List[List[float]] = []
    for index, vector in enumerate(vectors):
        if len(vector) != want:
            raise EmbeddingClientError(f"embedding {index} has dim {len(vector)}; expected {want}")
        out.append([float(v) for v in vector])
    return out


def embed_texts(texts: Sequence[str]) -> EmbeddingResult:
    if not embeddings_enabled():
        raise EmbeddingClientError("EMBEDDINGS_BASE_URL is not configured")
    cleaned = _validate_texts(texts)
    if _embedder is not None:
        result = _embedder(cleaned)
        if len(result.vectors) != len(cleaned):
            raise EmbeddingClientError(
                f"embedding response count mismatch: got {len(result.vectors)}, expected {len(cleaned)}")
        return EmbeddingResult(vectors=_validate_vectors(result.vectors), model=result.model)

    if httpx is None:
        raise EmbeddingClientError("embedding HTTP request failed: httpx is not installed")
    payload = {"texts": cleaned, "model": settings.embeddings_model_id}
    url = f"{_normalize_base_url(settings.embeddings_base_url)}/embed"
    try:
        with httpx.Client(timeout=httpx.Timeout(settings.embeddings_timeout_s)) as client:
            response = client.post(url, json=payload)
    except httpx.HTTPError as exc:
        raise EmbeddingClientError(f"embedding HTTP request failed: {exc}") from exc
    if response.status_code != 200:
        detail = response.text.strip()[:400]
        raise EmbeddingClientError(f"embedding service returned {response.status_code}: {detail}")
    body = response.json()
    raw = body.get("embeddings")
    if not isinstance(raw, list):
        raise EmbeddingClientError("embedding response missing 'embeddings' list")
    if len(raw) != len(cleaned):
        raise EmbeddingClientError(
            f"embedding response count mismatch: got {len(raw)}, expected {len(cleaned)}")
    return EmbeddingResult(vectors=_validate_vectors(raw),
                           model=str(body.get("model") or settings.embeddings_model_id))


def embed_texts_batched(texts: Sequence[str], batch_size: Optional[int] = None) -> EmbeddingResult:
    cleaned = _validate_texts(texts)
    size = batch_size or settings.embeddings_batch_size
    if size <= 0:
        raise EmbeddingClientError("batch size must be > 0")
    vectors: List[List[float]] = []
    model_used = settings.embeddings_model_id
    for start in range(0, len(cleaned), size):
        result = embed_texts(cleaned[start:start + size])
        vectors.extend(result.vectors)
        model_used = result.model
    return EmbeddingResult(vectors=vectors, model=model_used)


class SyntheticEmbedder:
    """Deterministic stand-in for the gateway: text -> row of the synthetic query stream
    (same generator as the corpus, seed = query seed).  The row index is a
    stable 63-bit hash of the text, so equal texts embed equally.  Vectors are produced on the
    GPU by the engine's generator and returned as Python floats like the HTTP client does."""

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
