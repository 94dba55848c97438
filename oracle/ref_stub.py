"""oracle/ref_stub.py -- TEST INFRASTRUCTURE ONLY (works only where /root/reference exists).

Imports the reference's *pure* functions live from /root/reference so that the golden-vector
generator (tests/golden/make_golden.py) and the CPU test tier can check our restatements
against the reference's own code rather than against another restatement.

The reference's app/retrieve.py imports sqlalchemy (absent in this image) and app/db.py
creates an engine at import time (app/db.py:11).  A six-line stand-in module is injected into
sys.modules for the duration of the import; nothing under /root/reference is modified or
copied.  /root/reference does not exist on the GPU box: every consumer must guard with
``available()`` and the GPU tier uses the committed fixtures in tests/golden/ instead.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CADENCE_REFERENCE_ROOT", "/root/reference")
_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "app", "retrieve.py"))


class _NoDB:
    def connect(self):
        raise RuntimeError("no database in this environment (oracle/ref_stub.py)")

    begin = connect


def load():
    """Returns a namespace with the reference's pure callables."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    if "sqlalchemy" not in sys.modules:
        sa = types.ModuleType("sqlalchemy")
        sa.text = lambda s: s
        sa.create_engine = lambda *a, **k: _NoDB()
        sys.modules["sqlalchemy"] = sa
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ref_retrieve = importlib.import_module("app.retrieve")
    ref_ingest = importlib.import_module("app.ingest")
    ref_embeddings = importlib.import_module("app.embeddings")
    ref_config = importlib.import_module("app.config")
    ref_schemas = importlib.import_module("app.schemas")
    ns = types.SimpleNamespace(
        retrieve=ref_retrieve,
        ingest=ref_ingest,
        embeddings=ref_embeddings,
        settings=ref_config.settings,
        schemas=ref_schemas,
        rrf_merge=ref_retrieve._rrf_merge,
        choose_dense_mode=ref_retrieve._choose_dense_mode,
        dense_has_scoping=ref_retrieve._dense_has_scoping,
        vector_literal=ref_retrieve._vector_literal,
        build_debug_lane=ref_retrieve._build_debug_lane,
        build_filter_clause=ref_retrieve._build_filter_clause,
        extract_tech_tokens=ref_ingest.extract_tech_tokens,
        RetrieveFilters=ref_schemas.RetrieveFilters,
    )
    _loaded = ns
    return ns
