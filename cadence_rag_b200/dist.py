"""Row-sharded multi-GPU search: one process per GPU, each owning a contiguous
row range of the corpus; the query batch is replicated; each rank's local top-k lists are
exchanged and merged with the global ordering rule, so every rank ends with the identical result.

Two transports for the exchange + merge (same bits):
  peer  one kernel per rank (K4p, csrc/peer.cu): push the local lists into every peer's CUDA-IPC mapped
        buffer over NVLink, wait for the peers' lists on local memory, merge.  Default on one node.
  nccl  ONE all-gather of a packed buffer (gloo in the CPU tests) + the K4 merge kernel.  Payload per
        rank: nq * (2k+1) * 8 bytes (scores f64 + ids i64 + count) -- 103 KB at nq=128, k=50;
        latency-bound, so it is a single collective per batch.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

from . import _ffi


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous rows of rank `rank`: [first, first + count). Rank r owns rows
    [r*ceil(N/R), min(N, (r+1)*ceil(N/R)))."""
    per = (n_total + world - 1) // world
    first = min(n_total, rank * per)
    return first, max(0, min(n_total, first + per) - first)


def pack_results(ids, scores, n):
    """(ids[nq,k] i64, scores[nq,k] f64, n[nq] i32) -> int64 [nq, 2k+1] (bit views, no rounding)."""
    import torch
    return torch.cat([scores.view(torch.int64), ids, n.to(torch.int64).unsqueeze(1)], dim=1).contiguous()


def unpack_results(packed, k: int):
    """int64 [R, nq, 2k+1] -> (scores[R,nq,k] f64, ids[R,nq,k] i64, n[R,nq] i32)."""
    import torch
    scores = packed[..., :k].contiguous().view(torch.float64)
    ids = packed[..., k:2 * k].contiguous()
    n = packed[..., 2 * k].to(torch.int32).contiguous()
    return scores, ids, n


def gather_shard_results(ids, scores, n, group=None):
    """All-gather every rank's local lists: returns (scores[R,nq,k], ids[R,nq,k], n[R,nq]) on each
    rank, rank-major.  Works on CUDA tensors (NCCL) and CPU tensors (gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    k = ids.shape[1]
    mine = pack_results(ids, scores, n)
    # concatenated-along-dim-0 output: the one layout both NCCL and gloo accept
    out = torch.empty((world * mine.shape[0], mine.shape[1]), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return unpack_results(out.view(world, mine.shape[0], mine.shape[1]), k)


def merge_shard_results(scores, ids, n, k: int, stream=None):
    """K4 on the device: [R,nq,k] lists -> (ids[nq,k], scores[nq,k], n[nq]), order (score desc,
    NaN last, id asc).  CUDA tensors only -- there is no CPU merge in the product."""
    import torch
    if not scores.is_cuda:
        raise _ffi.DenseEngineError("merge_shard_results needs CUDA tensors (no CPU fallback)",
                                    _ffi.CDR_ERR_NO_DEVICE)
    R, nq, kk = scores.shape
    assert kk == k and ids.shape == scores.shape and tuple(n.shape) == (R, nq)
    scores = scores.contiguous(); ids = ids.contiguous(); n = n.to(torch.int32).contiguous()
    dev = scores.device
    out_sc = torch.empty((nq, k), dtype=torch.float64, device=dev)
    out_id = torch.empty((nq, k), dtype=torch.int64, device=dev)
    out_n = torch.empty((nq,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _ffi.check(_ffi.lib().cdr_topk_merge(_ffi.ptr(scores), _ffi.ptr(ids), _ffi.ptr(n), R, nq, k,
                                             _ffi.ptr(out_sc), _ffi.ptr(out_id), _ffi.ptr(out_n),
                                             _ffi.stream_ptr(stream)), "cdr_topk_merge")
    return out_id, out_sc, out_n


class PeerExchange:
    """NVLink peer-memory transport of the exchange step (csrc/peer.cu, K4p): every rank's receive
    buffer is mapped into its peers with CUDA IPC and ONE kernel per rank pushes the local lists to
    all peers, waits for theirs and merges -- no NCCL call and no pack/unpack copies per batch.
    Construction is collective (the IPC handles travel through ``all_gather_object``)."""

    def __init__(self, device: int, group=None, max_nq: int = 1024, max_k: int = 64):
        import torch
        import torch.distributed as dist
        self.device = device
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.max_k = max_k
        self._h = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(_ffi.CDR_PEER_HANDLE_BYTES)
        with torch.cuda.device(device):
            _ffi.check(_ffi.lib().cdr_peer_group_create(ctypes.byref(self._h), device, self.rank, self.world,
                                                        max_nq, max_k, handle), "cdr_peer_group_create")
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        status = _ffi.lib().cdr_peer_group_connect(self._h, b"".join(handles))
        # every rank must take the same transport: agree on the outcome (a CPU tensor under gloo: two ranks may share
        # ONE device there -- CUDA IPC maps a buffer of the same device as well -- which is how the 1-GPU test tier
        # exercises the multi-rank exchange)
        on_cpu = dist.get_backend(group) != "nccl"
        ok = torch.tensor([1 if status == _ffi.CDR_OK else 0], dtype=torch.int32,
                          device="cpu" if on_cpu else f"cuda:{device}")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            msg = _ffi.last_error() if status != _ffi.CDR_OK else "a peer rank could not map the buffers"
            self.close()
            raise _ffi.DenseEngineError(f"peer-memory exchange unavailable: {msg}", _ffi.CDR_ERR_UNSUPPORTED)

    def close(self) -> None:
        if self._h:
            _ffi.lib().cdr_peer_group_destroy(self._h)
            self._h = ctypes.c_void_p()

    def exchange_merge(self, ids, scores, n, k: int, stream=None):
        """Local (ids[nq,k], scores[nq,k], n[nq]) CUDA tensors -> the global top-k on every rank."""
        import torch
        nq = int(ids.shape[0])
        dev = ids.device
        out_sc = torch.empty((nq, k), dtype=torch.float64, device=dev)
        out_id = torch.empty((nq, k), dtype=torch.int64, device=dev)
        out_n = torch.empty((nq,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _ffi.check(_ffi.lib().cdr_peer_exchange_merge(self._h, _ffi.ptr(scores), _ffi.ptr(ids), _ffi.ptr(n), nq, k,
                                                          _ffi.ptr(out_sc), _ffi.ptr(out_id), _ffi.ptr(out_n),
                                                          _ffi.stream_ptr(stream)), "cdr_peer_exchange_merge")
        return out_id, out_sc, out_n


class ShardedSearcher:
    """One rank's view of a row-sharded table.

    transport: "peer" = fused push + merge kernel over NVLink peer memory (K4p), "nccl" = one
    all-gather + the K4 merge kernel, "auto" (default; env CADENCE_EXCHANGE overrides) = peer when the
    ranks can map each other's memory (one node, CUDA tensors), else nccl.  Both give identical bits."""

    def __init__(self, store, group=None, transport: str = "auto", max_nq: int = 1024, max_k: int = 64):
        import os
        import torch.distributed as dist
        self.store = store
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        transport = os.environ.get("CADENCE_EXCHANGE", transport)
        if transport not in ("auto", "peer", "nccl"):
            raise ValueError(f"transport {transport!r}: expected auto, peer or nccl")
        self.peer: Optional[PeerExchange] = None
        self.transport = "none" if self.world == 1 else "nccl"
        if transport == "peer" and self.world > 1 and dist.get_backend(group) != "nccl":
            import torch
            if not torch.cuda.is_available():
                raise _ffi.DenseEngineError("peer transport needs CUDA ranks", _ffi.CDR_ERR_UNSUPPORTED)
        if self.world > 1 and transport in ("auto", "peer") and (dist.get_backend(group) == "nccl" or transport == "peer"):
            try:
                self.peer = PeerExchange(store.device, group, max_nq=max_nq, max_k=max_k)
                self.transport = "peer"
            except _ffi.DenseEngineError:
                if transport == "peer":
                    raise

    def close(self) -> None:
        if self.peer is not None:
            self.peer.close()
            self.peer = None

    def search(self, queries_dev, k: int, allow=None, mode: str = "exact", shared: bool = False):
        """queries_dev: [nq, dim] CUDA tensor replicated on every rank.  Returns the global
        (ids, scores, n) on every rank.  mode: "exact" (fp32 scan; shared: see DenseStore.search_exact), "scan_bf16"
        (single-query ann lane) or "batch" (bf16 tensor-core lane)."""
        if self.peer is not None and k <= self.peer.max_k and getattr(queries_dev, "is_cuda", False):
            return self._search_one_call(queries_dev, k, allow, mode, shared)
        if mode == "exact":
            ids, scores, n = self.store.search_exact(queries_dev, k, allow, shared=shared)
        elif mode == "scan_bf16":           # mode "ann" for single queries: one scan of the bf16 rows per query
            ids, scores, n = self.store.search_scan_bf16(queries_dev, k, allow)
        else:
            ids, scores, n = self.store.search_batch(queries_dev, k, allow)
        if self.world == 1:
            return ids, scores, n
        if self.peer is not None and k <= self.peer.max_k:
            return self.peer.exchange_merge(ids, scores, n, k)
        g_sc, g_id, g_n = gather_shard_results(ids, scores, n, self.group)
        return merge_shard_results(g_sc, g_id, g_n, k)

    def _search_one_call(self, queries_dev, k: int, allow, mode: str, shared: bool):
        """Peer transport: the local lane + the K4p exchange + merge through ONE C call (``cdr_search_sharded``), so
        the step's launches are enqueued back to back instead of across two trips through Python."""
        import torch
        lane = {"exact": _ffi.CDR_DENSE_LANE_EXACT_F32_SHARED if shared else _ffi.CDR_DENSE_LANE_EXACT_F32,
                "scan_bf16": _ffi.CDR_DENSE_LANE_SCAN_BF16}.get(mode, _ffi.CDR_DENSE_LANE_BATCH_BF16)
        q = queries_dev
        if q.dtype is not torch.float32 or not q.is_contiguous():
            q = q.to(dtype=torch.float32).contiguous()
        if q.dim() == 1:
            q = q.unsqueeze(0)
        if int(q.shape[1]) != self.store.dim:
            raise _ffi.DenseEngineError(f"query dim {int(q.shape[1])} != store dim {self.store.dim}")
        nq, dev = int(q.shape[0]), q.device
        # one allocation for the three outputs (a single-query step is a few tens of microseconds of kernels: every
        # allocator round trip on the host shows up as GPU idle time in its latency)
        buf = torch.empty((2 * nq * k + (nq + 1) // 2,), dtype=torch.int64, device=dev)
        base = buf.data_ptr()              # [scores f64 | ids i64 | counts i32]; the views are built after the launches
        _ffi.check(_ffi.lib().cdr_search_sharded(self.store.handle, self.peer._h, lane, _ffi.ptr(q), nq, k, _ffi.ptr(allow),
                                                 base, base + nq * k * 8, base + 2 * nq * k * 8,
                                                 _ffi.current_stream_ptr(dev.index)), "cdr_search_sharded")
        out_sc = buf[:nq * k].view(torch.float64).view(nq, k)
        out_id = buf[nq * k:2 * nq * k].view(nq, k)
        out_n = buf[2 * nq * k:].view(torch.int32)[:nq]
        return out_id, out_sc, out_n
