"""Settings kept from the reference's app/config.py:10-16 (the seven EMBEDDINGS_* flags of the
dense lane), read from the environment with the same names (case-insensitive, no prefix), plus
CADENCE_GPU_* knobs of this engine.  Tests mutate the singleton with monkeypatch.setattr exactly
like the reference's tests do (tests/unit/test_retrieve_planner.py:16)."""
from __future__ import annotations

import os
from dataclasses import dataclass, fields


def _env(name: str, default, cast):
    for key in (name.upper(), name.lower()):
        if key in os.environ:
            try:
                return cast(os.environ[key])
            except ValueError:
                return default
    return default


@dataclass
class Settings:
    # reference: app/config.py:10-16
    embeddings_base_url: str = ""
    embeddings_model_id: str = "Qwen/Qwen3-Embedding-4B"
    embeddings_dim: int = 1024
    embeddings_timeout_s: float = 180.0
    embeddings_batch_size: int = 32
    embeddings_exact_scan_threshold: int = 2000
    embeddings_hnsw_ef_search: int = 80
    # this engine
    cadence_gpu_device: int = 0
    # mode "ann" is served by the batched bf16 tensor-core lane when at least this many queries are in flight (and
    # the lane is the faster one for the table); smaller batches use an HBM-bound scan.  3: over 1 M rows one request scans
    # the bf16 rows in 0.30 ms, two share one such pass (0.39 ms), and from three on one step of the tensor-core lane
    # (0.44 ms for any batch up to 64) beats a pair plus a single pass (0.65 ms) -- profiles/r02/bf16_share_probe.jsonl
    cadence_gpu_ann_min_batch: int = 3
    # 1: single requests whose planner mode is "ann" scan the bf16 copy of the rows (half the bytes,
    # candidate lists twice as wide, exact re-score: recall ~1.0) instead of the fp32 rows; 0: always the exact scan
    cadence_gpu_ann_bf16_scan: int = 1

    @classmethod
    def from_env(cls) -> "Settings":
        s = cls()
        for f in fields(cls):
            cast = {str: str, int: int, float: float}[type(getattr(s, f.name))]
            setattr(s, f.name, _env(f.name, getattr(s, f.name), cast))
        return s


settings = Settings.from_env()
