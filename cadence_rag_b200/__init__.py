"""cadence_rag_b200 -- B200-native dense-retrieval engine behind cadence-rag's /retrieve dense lane.

Python host code (this package) mirrors the call surface of the reference's
``app/retrieve.py`` and ``app/embeddings.py`` and calls hand-written sm_100a CUDA kernels
through the C ABI declared in ``include/cadence_dense.h`` (``libcadence_dense.so``).
PyTorch is used only to own device buffers and streams.  There is no CPU fallback: without
the built library or without a CUDA device every compute entry point raises
:class:`DenseEngineError`.
"""
from ._ffi import DenseEngineError, abi_version, kernel_launch_count, library_path  # noqa: F401
from .config import settings  # noqa: F401

__all__ = ["DenseEngineError", "abi_version", "kernel_launch_count", "library_path", "settings"]
