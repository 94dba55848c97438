#!/usr/bin/env python
"""bench.py -- queries/sec of the /retrieve dense lane (BASELINE.json configs[1]):
1M x 1024 fp32 corpus, single-query exact cosine scan + top-k=50 (HBM-bound GEMV), on N B200s.

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the reference's CPU path (pgvector restated) timed alone

A "step" is one batch of Q distinct single-query scans (Q = --queries-per-step, default 64) over
the resident corpus: every query reads every row once (no cross-query reuse: the corpus is far
larger than L2), keeps its top-(k+14) in the scan kernel, is re-scored in fp64 and ordered.  With
N > 1 the 1M-row corpus is row-sharded (strong scaling: total work fixed), each rank scans its
shard, and the per-rank top-k lists are exchanged and merged on every rank -- by the K4p peer-memory
kernel over NVLink (default), or ONE NCCL all-gather + the K4 kernel (CADENCE_EXCHANGE=nccl).

Other lines: --workload batch_bf16 (configs[2]; under torchrun configs[4]), --workload hybrid
(configs[3] + configs[0]).  The default line also carries `exact_batch_shared_reads` (the same steps with
3 queries sharing every streamed tile; informational) and the CPU baselines (exact scan + HNSW, restated).

value    : whole-job queries/sec with the queries already resident in HBM (device-timed).
e2e      : the same through the facade with HOST buffers (numpy in, numpy out): H2D of the
           queries and D2H of ids/scores/counts inside the timed region, every step.
roofline : dominant kernel = exact_scan_kernel; achieved = rows_local*dim*4 bytes per query /
           its CUDA-event duration measured live on the launching stream.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "queries/sec (top-k=50, 1024-d) exact fp32 scan"
UNIT = "queries/s"
N_ROWS = 1_000_000
DIM = 1024
TOPK = 50


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries-per-step", type=int, default=64)
    ap.add_argument("--rows", type=int, default=N_ROWS)
    ap.add_argument("--cpu-sample-rows", type=int, default=200_000)
    ap.add_argument("--cpu-sample-queries", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch-queries", type=int, default=1024, help="batch_bf16: queries per batch (configs[2]: 1024)")
    ap.add_argument("--hnsw-rows", type=int, default=10_000, help="batch_bf16: rows of the CPU HNSW baseline's sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--bf16-only", action="store_true", help="batch_bf16: keep only bf16 rows resident (C5 residency)")
    ap.add_argument("--workload", default="exact_f32", choices=["exact_f32", "batch_bf16", "hybrid", "ingest"],
                    help="exact_f32 = BASELINE configs[1] (default, the contract line); "
                         "batch_bf16 = configs[2]: 10M x 1024 bf16, 1024 queries on the tcgen05 lane")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """nvidia-smi sampled every 50 ms in a side process.  Started BEFORE the warm-up (the tool needs ~0.2 s to
    deliver its first sample); mark_begin()/mark_end() bracket the timed region and stop() reports the samples
    that arrived inside it.  A timed region shorter than the sampling period (8 GPUs, short steps) has none:
    the caller then keeps the same workload running untimed (`extend`) until a few samples exist."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.begin = 0
        self.end = None
        self.extended = False

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark_begin(self):
        self.begin = len(self.lines)

    def mark_end(self):
        self.end = len(self.lines)

    def samples_in_region(self) -> int:
        return (len(self.lines) if self.end is None else self.end) - self.begin

    def extend(self, step_fn, min_samples: int = 3, max_seconds: float = 1.5):
        """Keep the workload running (untimed) until the region holds min_samples samples."""
        if not self.proc or self.samples_in_region() >= min_samples:
            return
        self.extended = True
        t0 = time.perf_counter()
        while len(self.lines) - self.begin < min_samples and time.perf_counter() - t0 < max_seconds:
            step_fn()
        self.end = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, power = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        region = self.lines[self.begin:self.end] if self.end is not None else self.lines[self.begin:]
        for ln in region:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(power),
               "samples": len(sm), "reasons": sorted(reasons)}
        if self.extended:
            out["note"] = "timed region shorter than the sampling period: sampled over the same steps run on, untimed"
        return out


def workload_config(args, world):
    label = "BASELINE configs[1]" if args.rows == N_ROWS else "north_star headline (configs[1] kernel)"
    return {"workload": f"{label}: {args.rows} x {DIM} fp32 corpus, single-query exact cosine scan + "
                        f"top-k={TOPK}; step = {args.queries_per_step} distinct queries, one scan of the corpus per query",
            "rows": args.rows, "dim": DIM, "k": TOPK, "queries_per_step": args.queries_per_step}


# ----------------------------------------------------------------------------- CPU baseline
def cpu_baseline(rows_total: int, sample_rows: int, sample_queries: int, threads: int, target_seconds: float = 12.0):
    """pgvector-restated exact scan (oracle/pgvector_restated.c) on a bounded sample, scaled to the
    workload's row count (the scan is linear in rows).  The sample is `sample_rows` rows (800 MB at
    200 000 x 1024 fp32: larger than the host's last-level cache, so the scan streams from DRAM as
    it would over the full corpus) scanned by distinct queries until `target_seconds` of CPU work
    have elapsed (at least `sample_queries` queries)."""
    import numpy as np
    from oracle import cpu_oracle as orc
    sample_rows = min(sample_rows, rows_total)
    if threads <= 0:
        threads = os.cpu_count() or 1      # explicit: launchers may export OMP_NUM_THREADS=1
    x = orc.synth_rows(20260209, 0, sample_rows)
    pool = 256
    qs = orc.synth_rows(20260210, 10_000_000, pool)
    orc.exact_scan(qs[0], x, TOPK, variant=orc.VARIANT_PGV32, nthreads=threads)   # warm
    done = 0
    t0 = time.perf_counter()
    while True:
        orc.exact_scan(qs[done % pool], x, TOPK, variant=orc.VARIANT_PGV32, nthreads=threads)
        done += 1
        dt = time.perf_counter() - t0
        if done >= sample_queries and dt >= target_seconds:
            break
        if done >= 100_000:
            break
    cores = threads
    qps_sample = done / dt
    return {"value": qps_sample * sample_rows / rows_total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{done} queries x {sample_rows} of {rows_total} rows (pgvector 0.8.1 cosine loop "
                      f"restated in C, fp32 accumulate, {cores} OpenMP threads; rate scaled by rows)",
            "seconds": dt}


def cpu_baseline_hnsw(rows_sample: int, threads: int, target_seconds: float = 8.0):
    """The reference's mode "ann" restated on the CPU (oracle/hnsw_baseline.cc: pgvector's HNSW parameters
    m=16, ef_construction=64, ef_search=80, cosine via normalised inner product) over a bounded sample
    of the synthetic corpus: build excluded, queries on all host cores (one query per thread, like
    concurrent backends) and on one thread, recall@50 against the exact oracle.  Graph-walk cost grows
    ~log(rows), so the rate is reported for the sample size as is, NOT scaled; "restated, not Postgres"."""
    import numpy as np
    from oracle import cpu_oracle as orc
    if threads <= 0:
        threads = os.cpu_count() or 1
    x = orc.synth_rows(20260209, 0, rows_sample)
    qs = orc.synth_rows(20260210, 20_000_000, 1024)
    t0 = time.perf_counter()
    index = orc.HnswBaseline(x)
    build_s = time.perf_counter() - t0
    index.search(qs[:64], TOPK, 80, nthreads=threads)
    done, t0 = 0, time.perf_counter()
    while True:
        rows, _sims, _n = index.search(qs, TOPK, 80, nthreads=threads)
        done += qs.shape[0]
        dt = time.perf_counter() - t0
        if dt >= target_seconds:
            break
    t1 = time.perf_counter()
    index.search(qs[:128], TOPK, 80, nthreads=1)
    one = 128 / (time.perf_counter() - t1)
    recall = float(np.mean([len(set(rows[i].tolist()) & set((orc.exact_scan(qs[i], x, TOPK)[0] - 1).tolist())) / TOPK
                            for i in range(64)]))
    index.close()
    return {"value": done / dt, "unit": UNIT, "cores": threads, "kind": "port", "single_thread_value": one,
            "recall_at_50_vs_exact": recall, "rows": rows_sample, "m": 16, "ef_construction": 64, "ef_search": 80,
            "build_seconds": build_s,
            "sample": f"{done} queries over a {rows_sample}-row sample of the synthetic corpus (HNSW restated, not Postgres; "
                      "rate not scaled to the full corpus)"}


def run_reference(args):
    """--impl reference: the reference's own implementation of the path is SQL on Postgres+pgvector,
    which cannot run in this image (no postgres/pgvector/sqlalchemy); the arm times the C
    restatement of pgvector's exact scan on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import cpu_oracle as orc
    import numpy as np
    rows = min(args.cpu_sample_rows, args.rows)
    cores = os.cpu_count() or 1            # explicit: torchrun exports OMP_NUM_THREADS=1 to its workers
    q_per_step = max(1, args.queries_per_step)      # same step as the B200 arm: Q distinct single-query scans
    x = orc.synth_rows(20260209, 0, rows)
    n_q = (args.steps + args.warmup) * q_per_step
    qs = orc.synth_rows(20260210, 0, min(n_q, 4096))
    qi = 0
    for _ in range(args.warmup):
        for _ in range(q_per_step):
            orc.exact_scan(qs[qi % 4096], x, TOPK, variant=orc.VARIANT_PGV32, nthreads=cores); qi += 1
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(q_per_step):
            orc.exact_scan(qs[qi % 4096], x, TOPK, variant=orc.VARIANT_PGV32, nthreads=cores); qi += 1
    dt = time.perf_counter() - t0
    value = args.steps * q_per_step / dt * rows / args.rows
    sample = (f"{q_per_step} queries/step x {rows} of {args.rows} rows, rate scaled by rows; pgvector 0.8.1 "
              f"exact-scan loop restated in C (oracle/pgvector_restated.c), {cores} OpenMP threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


# ----------------------------------------------------------------------------- B200 arm
def run_batch_bf16(args):
    """Secondary line (BASELINE configs[2] on 1 GPU, configs[4] row-sharded on N GPUs): rows x 1024
    bf16 corpus, 1024 queries per step on the tcgen05 lane (K2) + exact re-score; with N > 1 the
    corpus is row-sharded and the per-rank top-k lists are all-gathered and merged (K4).
    roofline: tensor-bound, 2*nq*rows_local*dim FLOP per step per GPU."""
    import ctypes
    import numpy as np
    import torch
    import torch.distributed as dist
    from cadence_rag_b200 import _ffi
    from cadence_rag_b200.dist import ShardedSearcher, shard_range
    from cadence_rag_b200.store import DenseStore, SYNTH_QUERY_SEED, synth_rows_device
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    rows = args.rows if args.rows != N_ROWS else 10_000_000
    nq = args.batch_queries
    first, count = shard_range(rows, rank, world)
    store = DenseStore("chunks", count, dim=DIM, device=local_rank, fp32=not args.bf16_only, bf16=True)
    store.append_synthetic(count, first_row=first)
    store.finalize()
    searcher = ShardedSearcher(store)
    total = args.warmup + args.steps
    q_dev = synth_rows_device(SYNTH_QUERY_SEED, 0, total * nq, DIM, device=local_rank).view(total, nq, DIM)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for s in range(args.warmup):
        searcher.search(q_dev[s], TOPK, mode="ann")
    barrier()
    _ffi.lib().cdr_prof_enable(1)
    launches0 = _ffi.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    ev0.record()
    for s in range(args.warmup, total):
        out = searcher.search(q_dev[s], TOPK, mode="ann")
    ev1.record()
    barrier()
    sampler.mark_end()
    ms = ev0.elapsed_time(ev1)
    k_ms, k_n = ctypes.c_double(0), ctypes.c_int64(0)
    _ffi.check(_ffi.lib().cdr_prof_read(1, ctypes.byref(k_ms), ctypes.byref(k_n)))
    per_launch = np.zeros(int(k_n.value), dtype=np.float64)
    got_n = ctypes.c_int64(0)
    _ffi.check(_ffi.lib().cdr_prof_read_launches(1, _ffi.ptr(per_launch), per_launch.size, ctypes.byref(got_n)))
    segs = max(1, int(k_n.value) // max(args.steps, 1))
    last_step_launch_ms = [round(float(v), 4) for v in per_launch[-segs:]]
    _ffi.lib().cdr_prof_enable(0)
    launches = _ffi.kernel_launch_count() - launches0
    # too few clock samples (short region): every rank keeps stepping, untimed, while rank 0 samples
    # (a step COUNT agreed by all ranks: every rank must issue the same sequence of exchanges)
    n_ext = min(20000, int(0.6 / max(ms / 1e3 / max(args.steps, 1), 1e-5)) + 1) if (rank == 0 and sampler.samples_in_region() < 3 and sampler.proc) else 0
    need_more = torch.tensor([n_ext], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.broadcast(need_more, 0)
    if int(need_more.item()):
        for _ in range(int(need_more.item())):
            searcher.search(q_dev[total - 1], TOPK, mode="ann")
        barrier()
        if rank == 0:
            sampler.extended = True
            sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, float(launches), k_ms.value], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, launches, gemm_ms = float(tmax[0]), int(tsum[1]), float(tmax[2])
    else:
        gemm_ms = k_ms.value
    # recall@50 of the last batch against the exact fp32 lane (same sharded corpus), 64 queries
    recall = None
    if store.has_fp32:
        nr = min(64, nq)
        e_ids, _, _ = searcher.search(q_dev[total - 1][:nr].contiguous(), TOPK, mode="exact")
        got = out[0][:nr].cpu().numpy(); want = e_ids.cpu().numpy()
        recall = float(np.mean([len(set(got[i]) & set(want[i])) / TOPK for i in range(nr)]))
    # the exact fp32 lane over the same (sharded) corpus, one query per request: device-timed latency
    exact_q1 = None
    if store.has_fp32:
        for i in range(3):
            searcher.search(q_dev[0][i % nq:i % nq + 1].contiguous(), TOPK, mode="exact")
        barrier()
        xa, xb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        xa.record()
        for i in range(20):
            searcher.search(q_dev[1][i % nq:i % nq + 1].contiguous(), TOPK, mode="exact")
        xb.record()
        barrier()
        xt = torch.tensor([xa.elapsed_time(xb) / 20], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(xt, op=dist.ReduceOp.MAX)
        exact_q1 = {"ms_per_query": float(xt[0]), "hbm_gbs_per_gpu": float(count) * DIM * 4 / (float(xt[0]) / 1e3) / 1e9}
    # mode "ann" for ONE query: the same scan over the bf16 rows (half the bytes), exact re-score
    scan_q1 = None
    if store.has_bf16:
        for i in range(3):
            searcher.search(q_dev[0][i % nq:i % nq + 1].contiguous(), TOPK, mode="scan_bf16")
        barrier()
        xa, xb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        xa.record()
        for i in range(20):
            got1 = searcher.search(q_dev[1][i % nq:i % nq + 1].contiguous(), TOPK, mode="scan_bf16")
        xb.record()
        barrier()
        xt = torch.tensor([xa.elapsed_time(xb) / 20], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(xt, op=dist.ReduceOp.MAX)
        scan_q1 = {"ms_per_query": float(xt[0]), "hbm_gbs_per_gpu": float(count) * DIM * 2 / (float(xt[0]) / 1e3) / 1e9}
        if store.has_fp32:
            want1 = searcher.search(q_dev[1][19 % nq:19 % nq + 1].contiguous(), TOPK, mode="exact")
            scan_q1["recall_at_50_vs_exact_fp32_lane_last_query"] = len(set(got1[0][0].tolist()) & set(want1[0][0].tolist())) / TOPK
    # e2e with host buffers (pinned): H2D of the queries + D2H of the merged result every step
    q_pinned = torch.empty(q_dev.shape, dtype=torch.float32, pin_memory=True)
    q_pinned.copy_(q_dev); torch.cuda.synchronize()
    q_host = q_pinned.numpy()

    def e2e_step(s):
        if world == 1:
            return store.search_batch(q_host[s], TOPK)
        qd = torch.from_numpy(q_host[s]).cuda()
        ids, sc, n = searcher.search(qd, TOPK, mode="ann")
        return ids.cpu().numpy(), sc.cpu().numpy(), n.cpu().numpy()
    tt = None
    if not args.no_e2e:
        e2e_step(0)
        barrier()
        t0 = time.perf_counter()
        for s in range(args.warmup, total):
            e2e_step(s)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peaks = json.load(open(pk)) if os.path.exists(pk) else {}
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        flops_step_gpu = 2.0 * nq * count * DIM
        # DRAM bytes of all gemm_topk launches of one step, from an ncu capture of this command
        # (profiles/k2_traffic.json: bytes per corpus row), scaled to this shard
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "k2_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("dram_bytes_per_row") and tj.get("queries") == nq:
                traffic = tj["dram_bytes_per_row"] * count
        achieved = flops_step_gpu * args.steps / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        line = {"metric": "queries/sec (top-k=50, 1024-d) batched bf16 tcgen05 lane",
                "value": args.steps * nq / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"BASELINE configs[{2 if world == 1 else 4}]: {rows} x {DIM} bf16 corpus"
                                       f"{' row-sharded over %d GPUs' % world if world > 1 else ''}, batch {nq} queries, "
                                       f"tcgen05 GEMM with fused threshold top-k epilogue + exact re-score, top-k={TOPK}",
                           "rows": rows, "rows_per_gpu": count, "dim": DIM, "k": TOPK, "queries_per_step": nq,
                           "resident": "bf16 only" if args.bf16_only else "fp32 + bf16",
                           "l2": "inputs larger than L2", "recall_at_50_vs_exact_fp32_lane": recall,
                           "exchange": searcher.transport, "exact_fp32_lane_single_query": exact_q1,
                           "ann_bf16_scan_single_query": scan_q1},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": None if tt is None else {"value": args.steps * nq / float(tt[0]), "unit": UNIT,
                                                "h2d_bytes_per_step": nq * DIM * 4,
                                                "d2h_bytes_per_step": nq * TOPK * 16 + nq * 4},
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": achieved / peak if achieved else None, "traffic": traffic,
                             "algorithmic_flops_per_step": flops_step_gpu,
                             "corpus_stream_gbs": (float(count) * DIM * 2 * args.steps / (gemm_ms / 1e3) / 1e9) if gemm_ms > 0 else None,
                             "algorithmic_dram_bytes_per_step": float(count) * DIM * 2,
                             "kernel": "gemm_topk_kernel", "per": "GPU (max over ranks)",
                             "peak_source": "measured bf16_tflops_sustained (kernel timed inside a long step)",
                             "peak_burst": peaks.get("bf16_tflops"), "gemm_ms_per_step": gemm_ms / args.steps,
                             "launches_timed": int(k_n.value), "segment_launch_ms_last_step": last_step_launch_ms,
                             "gemm_ms_by_step": [round(float(per_launch[i * segs:(i + 1) * segs].sum()), 3)
                                                 for i in range(args.steps)]}}
        if world == 1 and not args.no_cpu_baseline:
            # the reference's two CPU paths for this config, timed on this box's host cores in the same run
            line["cpu_baseline"] = cpu_baseline(rows, args.cpu_sample_rows, args.cpu_sample_queries, 0)
            line["cpu_baseline_hnsw"] = cpu_baseline_hnsw(args.hnsw_rows, 0)
        emit(line)
    if world > 1:
        dist.barrier()
    searcher.close()
    store.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_hybrid(args):
    """Secondary line (BASELINE configs[3] + configs[0]): hybrid /retrieve over 1M chunks through the
    facade -- synthetic embedder, K6 filter bitmap + count, planner, K1 exact scan, host tech_tokens
    lane, K5 RRF -- for (a) a 10-call filter (2 000 candidate rows => mode "exact", the C1 shape) and
    (b) no filter.  Fused ranks are checked bit-exact against the restated pipeline on 8 queries."""
    import numpy as np
    import torch
    from cadence_rag_b200 import _ffi, embeddings, retrieve
    from cadence_rag_b200.config import settings
    from cadence_rag_b200.lexical import TechTokenIndex
    from cadence_rag_b200.retrieve import DenseEngine, RetrieveFilters
    from cadence_rag_b200.store import DenseStore, SYNTH_CORPUS_SEED, SYNTH_QUERY_SEED
    from oracle import cpu_oracle as orc
    from oracle import ports
    rows = args.rows
    torch.cuda.set_device(0)
    store = DenseStore("chunks", rows, dim=DIM, device=0, fp32=True, bf16=True)
    store.append_synthetic(rows)
    store.finalize()
    rng = np.random.default_rng(SYNTH_CORPUS_SEED)
    vocab = 10_000
    ntok = rng.integers(0, 4, size=rows)
    tok = np.minimum(rng.zipf(1.1, size=(rows, 3)) - 1, vocab - 1)
    index = TechTokenIndex()
    flat_rows = np.repeat(np.arange(rows), 3)[(np.arange(3)[None, :] < ntok[:, None]).reshape(-1)]
    flat_tok = tok.reshape(-1)[(np.arange(3)[None, :] < ntok[:, None]).reshape(-1)]
    order = np.lexsort((flat_rows, flat_tok))
    flat_rows, flat_tok = flat_rows[order], flat_tok[order]
    starts = np.searchsorted(flat_tok, np.arange(vocab + 1))
    for t in range(vocab):
        if starts[t + 1] > starts[t]:
            index.add_postings(f"TK-{t}", np.unique(flat_rows[starts[t]:starts[t + 1]]))
    eng = DenseEngine()
    eng.register(store, index)
    emb = embeddings.SyntheticEmbedder(seed=SYNTH_QUERY_SEED, dim=DIM)
    embeddings.set_embedder(emb)
    settings.embeddings_dim = DIM
    filt = RetrieveFilters(call_ids=list(range(10)))          # synthetic call id == call slot; 200 rows/call
    out = {}
    # parity of the fused ranks on a few queries (restated pipeline: oracle dense + port tech + port RRF)
    x_small = orc.synth_rows(SYNTH_CORPUS_SEED, 0, 2000)
    cols = store.host_columns()
    for qi in range(8):
        text = f"status of TK-{qi} and TK-{qi * 7 + 1} on v1.{qi}"
        got = retrieve.retrieve_ids(eng, text, filt, debug=True)
        q = np.array(emb([text]).vectors[0], dtype=np.float32)
        d_ids, _ = orc.exact_scan(q, x_small, TOPK)           # the 10 calls are rows 0..1999
        toks = retrieve.extract_tech_tokens(text)
        keep = cols["call_slot"] < 10
        hit_rows = np.unique(np.concatenate([index.postings(t) for t in toks] + [np.empty(0, dtype=np.int64)]))
        hit_rows = hit_rows[keep[hit_rows]]
        o = np.lexsort((cols["ids"][hit_rows], -cols["started_at"][hit_rows]))
        tech_ids = cols["ids"][hit_rows[o]][:50].tolist()
        want = ports.rrf_merge({"bm25": [], "tech_tokens": [{"chunk_id": i} for i in tech_ids],
                                "dense": [{"chunk_id": int(i)} for i in d_ids]}, "chunk_id")
        assert [(r["chunk_id"], sorted(h), s_) for r, h, s_ in want] == [tuple(t) for t in got["debug"]["fused"]["chunks"]], qi
        assert got["debug"]["dense"]["modes"]["chunks"] == "exact" and got["debug"]["dense"]["candidate_rows"]["chunks"] == 2000
    for name, f in (("filtered_10_calls_2000_rows", filt), ("unfiltered", None)):
        n = args.steps * 8
        for i in range(8):
            retrieve.retrieve_ids(eng, f"warm TK-{i} TK-{i + 3}", f)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            retrieve.retrieve_ids(eng, f"status of TK-{i % 500} and TK-{(i * 13) % 900}", f)
        dt = time.perf_counter() - t0
        out[name] = {"queries_per_s": n / dt, "ms_per_query": dt / n * 1e3, "queries": n}
        if f is None and settings.cadence_gpu_ann_bf16_scan:
            # unscoped requests plan "ann" and scan the bf16 rows by default; the same requests on the exact fp32 scan
            settings.cadence_gpu_ann_bf16_scan = 0
            try:
                for i in range(8):
                    retrieve.retrieve_ids(eng, f"warm TK-{i} TK-{i + 3}", f)
                t0 = time.perf_counter()
                for i in range(n):
                    retrieve.retrieve_ids(eng, f"status of TK-{i % 500} and TK-{(i * 13) % 900}", f)
                out[name]["queries_per_s_exact_fp32_scan"] = n / (time.perf_counter() - t0)
            finally:
                settings.cadence_gpu_ann_bf16_scan = 1
    # batched form: 64 requests per fused C call (cdr_hybrid_retrieve_host), embeddings and token ids prepared
    # outside the timed region (the embedder is a remote model in the reference), host buffers in and out
    dev_index = eng.device_tech_indexes["chunks"]
    B = 64
    texts = [f"status of TK-{i % 500} and TK-{(i * 13) % 900}" for i in range(B * 4)]
    qv = np.stack([np.asarray(emb([t]).vectors[0], dtype=np.float32) for t in texts])
    tok, nt = dev_index.encode_tokens([retrieve.extract_tech_tokens(t) for t in texts])
    for name, f in (("filtered_10_calls_2000_rows", filt), ("unfiltered", None)):
        spec = retrieve._filter_spec(store, f, f.call_ids if f else None)
        for b in range(2):
            store.hybrid_retrieve(qv[:B], TOPK, tech_index=dev_index, token_ids=tok[:B], n_tokens=nt[:B], filter_spec=spec)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = max(2, args.steps // 2)
        for r in range(reps):
            o = (r % 4) * B
            store.hybrid_retrieve(qv[o:o + B], TOPK, tech_index=dev_index, token_ids=tok[o:o + B], n_tokens=nt[o:o + B],
                                  filter_spec=spec)
        dt = time.perf_counter() - t0
        out[name]["batched_64_queries_per_s"] = reps * B / dt
        if f is None:
            # the same 64 unscoped requests (planner mode "ann") with the group's dense lane on the bf16 tensor cores
            ann = dict(spec, dense_lane=_ffi.CDR_DENSE_LANE_BATCH_BF16)
            exact_ids = store.hybrid_retrieve(qv[:B], TOPK, tech_index=dev_index, token_ids=tok[:B], n_tokens=nt[:B], filter_spec=spec)
            for b in range(2):
                got = store.hybrid_retrieve(qv[:B], TOPK, tech_index=dev_index, token_ids=tok[:B], n_tokens=nt[:B], filter_spec=ann)
            same = float(np.mean([len(set(got["dense_ids"][i].tolist()) & set(exact_ids["dense_ids"][i].tolist())) / TOPK for i in range(B)]))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for r in range(reps):
                o = (r % 4) * B
                store.hybrid_retrieve(qv[o:o + B], TOPK, tech_index=dev_index, token_ids=tok[o:o + B], n_tokens=nt[o:o + B],
                                      filter_spec=ann)
            dt = time.perf_counter() - t0
            out[name]["batched_64_ann_lane_queries_per_s"] = reps * B / dt
            out[name]["batched_64_ann_lane_recall_at_50_vs_exact_lane"] = same
    # concurrent clients: the reference serves /retrieve from a threadpool (app/main.py:184-186); 8 client threads,
    # each on its own CUDA stream, one request at a time per thread, through retrieve_ids
    import threading
    for name, f in (("filtered_10_calls_2000_rows", filt), ("unfiltered", None)):
        n_threads, per_thread = 8, max(8, args.steps)
        errs = []

        def client(t):
            try:
                with torch.cuda.stream(torch.cuda.Stream()):
                    for i in range(per_thread):
                        retrieve.retrieve_ids(eng, f"status of TK-{(t * 97 + i) % 500} and TK-{(i * 13 + t) % 900}", f)
            except Exception as exc:   # noqa: BLE001
                errs.append(repr(exc))
        for label, interval in (("concurrent_8_clients_queries_per_s", None), ("concurrent_8_clients_switchinterval_50us_queries_per_s", 5e-5)):
            # CPython hands the GIL over every 5 ms by default: a client that returns from the (GIL-free) C call
            # waits that long behind a peer running Python; sys.setswitchinterval(50 us) removes the convoy
            old_interval = sys.getswitchinterval()
            if interval is not None:
                sys.setswitchinterval(interval)
            threads = [threading.Thread(target=client, args=(t,)) for t in range(n_threads)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for th in threads:
                th.start()
            for th in threads:
                th.join()
            dt = time.perf_counter() - t0
            sys.setswitchinterval(old_interval)
            assert not errs, errs
            out[name][label] = n_threads * per_thread / dt
    # the same clients through the micro-batcher (one worker, one fused call per batch; requests with equal filters
    # form a group).  "distinct filters": every client scopes to its own 10-call window.
    for name, f_of in (("filtered_10_calls_2000_rows", lambda t: filt), ("unfiltered", lambda t: None),
                       ("distinct_filters_per_client", lambda t: RetrieveFilters(call_ids=list(range(10 * t, 10 * t + 10))))):
        for n_threads in (8, 32):
            per_thread = max(8, args.steps)
            batcher = retrieve.RequestBatcher(eng, max_batch=64, max_wait_s=2e-4)
            errs = []

            def bclient(t, f_of=f_of):
                try:
                    for i in range(per_thread):
                        batcher.retrieve_ids(f"status of TK-{(t * 97 + i) % 500} and TK-{(i * 13 + t) % 900}", f_of(t))
                except Exception as exc:   # noqa: BLE001
                    errs.append(repr(exc))
            threads = [threading.Thread(target=bclient, args=(t,)) for t in range(n_threads)]
            old_interval = sys.getswitchinterval()
            sys.setswitchinterval(5e-5)          # see above: the default 5 ms GIL hand-over makes thread timings erratic
            t0 = time.perf_counter()
            for th in threads:
                th.start()
            for th in threads:
                th.join()
            dt = time.perf_counter() - t0
            sys.setswitchinterval(old_interval)
            batcher.close()
            assert not errs, errs
            out.setdefault(name, {})[f"batcher_{n_threads}_clients_queries_per_s"] = n_threads * per_thread / dt
            out[name][f"batcher_{n_threads}_clients_mean_batch"] = batcher.requests_served / max(batcher.batches_served, 1)
    if os.environ.get("CADENCE_BENCH_HOST_PROFILE"):
        import cProfile, pstats
        pr = cProfile.Profile(); pr.enable()
        for i in range(300):
            retrieve.retrieve_ids(eng, f"status of TK-{i % 500} and TK-{(i * 13) % 900}", filt)
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(35)
    # C1: 2 000-row store, GPU exact-scan latency vs the CPU restatement (1 thread and all cores)
    small = DenseStore("chunks", 2000, dim=DIM, device=0, fp32=True, bf16=False)
    small.append_synthetic(2000); small.finalize()
    qs = orc.synth_rows(SYNTH_QUERY_SEED, 0, 64)
    qd = torch.from_numpy(qs).cuda()
    for i in range(5):
        small.search_exact(qd[i:i + 1], TOPK)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(64):
        small.search_exact(qd[i:i + 1], TOPK)
    ev1.record(); torch.cuda.synchronize()
    gpu_ms = ev0.elapsed_time(ev1) / 64
    t0 = time.perf_counter()
    for i in range(64):
        small.search_exact(qs[i], TOPK)
    host_ms = (time.perf_counter() - t0) / 64 * 1e3
    cpu = {}
    for th in (1, 0):
        t0 = time.perf_counter()
        for i in range(64):
            orc.exact_scan(qs[i], x_small, TOPK, variant=orc.VARIANT_PGV32, nthreads=th)
        cpu["1_thread" if th == 1 else f"{orc.num_threads()}_threads"] = (time.perf_counter() - t0) / 64 * 1e3
    line = {"metric": "hybrid /retrieve queries/sec (dense top-50 + tech_tokens lane + RRF k=60)",
            "value": out["unfiltered"]["queries_per_s"], "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": 8, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[3]: hybrid /retrieve, {rows} chunks, through retrieve_ids "
                                   "(host facade, one request at a time, one fused C call per table: K6 + tech lane + "
                                   "K1 + K5, one sync); batched_64 = 64 requests per fused call", "rows": rows, "k": TOPK},
            "hybrid": out, "fused_ranks_bit_exact_queries": 8,
            "c1_exact_scan_2000_rows": {"gpu_device_ms": gpu_ms, "gpu_host_buffers_ms": host_ms, "cpu_ms": cpu}}
    emit(line)
    embeddings.set_embedder(None)
    store.close(); small.close()
    return 0


def run_ingest(args):
    """Write side of the store: host rows -> resident rows (fp32 + inverse norm +
    normalised bf16), the backfill UPDATE in place, growth of a sealed store, snapshot save / load, building the
    device tech-token index, and the hierarchical (artifact -> shortlist -> scoped chunks) dense request."""
    import shutil
    import tempfile
    import numpy as np
    import torch
    from cadence_rag_b200 import retrieve
    from cadence_rag_b200.config import settings
    from cadence_rag_b200.lexical import DeviceTechIndex, TechTokenIndex
    from cadence_rag_b200.retrieve import DenseEngine
    from cadence_rag_b200.store import DenseStore, SYNTH_CORPUS_SEED, SYNTH_QUERY_SEED, synth_rows_device
    torch.cuda.set_device(0)
    settings.embeddings_dim = DIM
    rows = min(args.rows, 400_000)
    n_art = rows // 10
    out = {}
    x = synth_rows_device(SYNTH_CORPUS_SEED, 0, rows, DIM, device=0).cpu().numpy()        # host rows, as an ingest sees them
    ids = np.arange(1, rows + 1, dtype=np.int64)
    slots = [int(r) // 200 for r in range(rows)]

    def timed(fn, reps=1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps, r

    chunk = 65_536
    store = DenseStore("chunks", rows + chunk, dim=DIM, device=0)
    valid = np.ones(rows, dtype=bool); valid[::10] = False                                # every 10th row ingested with embedding NULL

    def load():
        for r0 in range(0, rows, chunk):
            r1 = min(rows, r0 + chunk)
            store.append(x[r0:r1], ids=ids[r0:r1], call_ids=slots[r0:r1], valid=valid[r0:r1])
    dt, _ = timed(load)
    out["append_host_rows"] = {"rows": rows, "rows_per_s": rows / dt, "gb_per_s": rows * DIM * 4 / dt / 1e9,
                               "note": "pageable numpy rows -> fp32 + inverse norm + bf16 resident, 65 536 rows per call"}
    dt, _ = timed(store.finalize)
    out["finalize_ms"] = dt * 1e3
    pend = store.pending_ids()
    m = int(pend.size)
    dt, _ = timed(lambda: [store.update_embeddings(pend[i:i + 4096], x[pend[i:i + 4096] - 1]) for i in range(0, m, 4096)])
    out["backfill_update_in_place"] = {"rows": m, "rows_per_s": m / dt, "note": "cdr_store_update_embeddings, 4 096 rows per call"}
    grow = synth_rows_device(SYNTH_CORPUS_SEED, rows, chunk, DIM, device=0).cpu().numpy()
    dt, _ = timed(lambda: store.append(grow, ids=np.arange(rows + 1, rows + chunk + 1, dtype=np.int64), call_ids=[rows // 200] * chunk))
    out["grow_sealed_store"] = {"rows": chunk, "rows_per_s": chunk / dt}
    tmp = tempfile.mkdtemp(prefix="cdr_snapshot_")
    try:
        dt, _ = timed(lambda: store.save(tmp))
        out["snapshot_save"] = {"rows": store.rows, "seconds": dt, "gb_per_s": store.rows * DIM * 4 / dt / 1e9}
        dt, restored = timed(lambda: DenseStore.load(tmp, device=0))
        out["snapshot_load"] = {"rows": restored.rows, "seconds": dt, "gb_per_s": restored.rows * DIM * 4 / dt / 1e9}
        q = synth_rows_device(SYNTH_QUERY_SEED, 0, 1, DIM, device=0)
        a, b = store.search_exact(q, TOPK), restored.search_exact(q, TOPK)
        torch.cuda.synchronize()
        assert torch.equal(a[0], b[0]) and torch.equal(a[1].view(torch.int64), b[1].view(torch.int64)), "restored store answers differently"
        restored.close()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    # device tech-token index (f-1): 0-3 tokens per row from a 10 000-token Zipf vocabulary
    rng = np.random.default_rng(1)
    n_rows = store.rows
    ntok = rng.integers(0, 4, size=n_rows)
    tok = np.minimum(rng.zipf(1.1, size=(n_rows, 3)) - 1, 9_999)
    mask = (np.arange(3)[None, :] < ntok[:, None]).reshape(-1)
    flat_rows, flat_tok = np.repeat(np.arange(n_rows), 3)[mask], tok.reshape(-1)[mask]
    order = np.lexsort((flat_rows, flat_tok))
    flat_rows, flat_tok = flat_rows[order], flat_tok[order]
    starts = np.searchsorted(flat_tok, np.arange(10_001))
    index = TechTokenIndex()
    for t in range(10_000):
        if starts[t + 1] > starts[t]:
            index.add_postings(f"TK-{t}", np.unique(flat_rows[starts[t]:starts[t + 1]]))
    dt, dev_index = timed(lambda: DeviceTechIndex(index, store))
    out["device_tech_index_build"] = {"rows": n_rows, "postings": int(flat_rows.size), "seconds": dt}
    # hierarchical dense request (f-4): artifacts -> call shortlist -> chunks scoped to the shortlist
    arts = DenseStore("artifact_chunks", n_art, dim=DIM, device=0)
    xa = synth_rows_device(SYNTH_CORPUS_SEED + 5, 0, n_art, DIM, device=0).cpu().numpy()
    arts.append(xa, ids=np.arange(1, n_art + 1, dtype=np.int64), call_ids=[int(r) // 20 for r in range(n_art)])
    arts.finalize()
    eng = DenseEngine()
    eng.register(store); eng.register(arts)
    qs = synth_rows_device(SYNTH_QUERY_SEED, 100, 64, DIM, device=0).cpu().numpy()
    with eng.connect() as conn:
        retrieve.fetch_chunks_dense_hierarchical(conn, qs[0], None, None)
        dt, res = timed(lambda: [retrieve.fetch_chunks_dense_hierarchical(conn, qs[i], None, None) for i in range(64)])
    out["hierarchical_dense_request"] = {"ms_per_request": dt / 64 * 1e3, "chunks": store.rows, "artifact_chunks": n_art,
                                         "shortlist_calls": len(res[-1]["call_shortlist"]),
                                         "scoped_candidate_rows": res[-1]["candidate_rows"]["chunks"], "modes": res[-1]["modes"]}
    line = {"metric": "store write side and f-rows (rows/s, seconds)", "value": out["append_host_rows"]["rows_per_s"], "unit": "rows/s",
            "n_gpus": 1, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"ingest of {rows} x {DIM} fp32 host rows + backfill / growth / snapshot / tech index / hierarchical request"},
            "ingest": out}
    emit(line)
    dev_index.close(); store.close(); arts.close()
    return 0


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries write to fd 1 behind Python's back (NCCL prints its version banner there on the first
    communicator): point fd 1 at stderr for the run, so that stdout carries the JSON line and nothing else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    """The run's ONE JSON line, on the real stdout."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "batch_bf16":
        return run_batch_bf16(args)
    if args.workload == "hybrid":
        return run_hybrid(args)
    if args.workload == "ingest":
        return run_ingest(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from cadence_rag_b200 import _ffi
    from cadence_rag_b200.dist import ShardedSearcher, shard_range
    from cadence_rag_b200.store import DenseStore, SYNTH_QUERY_SEED, synth_rows_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    _ffi.require_device()

    Q = args.queries_per_step
    first, count = shard_range(args.rows, rank, world)
    store = DenseStore("chunks", max(count, 1), dim=DIM, device=local_rank, fp32=True, bf16=False)
    store.append_synthetic(count, first_row=first)
    store.finalize()
    searcher = ShardedSearcher(store)
    total_steps = args.warmup + args.steps
    # distinct queries for every step, resident in HBM before the timed region
    q_dev = synth_rows_device(SYNTH_QUERY_SEED, 0, total_steps * Q, DIM, device=local_rank).view(total_steps, Q, DIM)
    q_pinned = torch.empty(q_dev.shape, dtype=torch.float32, pin_memory=True)   # e2e inputs: pinned host memory
    q_pinned.copy_(q_dev)
    q_host = q_pinned.numpy()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (the clock sampler starts first: nvidia-smi needs ~0.2 s to deliver its first sample)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for s in range(args.warmup):
        searcher.search(q_dev[s], TOPK)
    barrier()

    # ---- timed: device-resident queries
    _ffi.lib().cdr_prof_enable(1)
    launches0 = _ffi.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    ev0.record()
    last = None
    for s in range(args.warmup, total_steps):
        last = searcher.search(q_dev[s], TOPK)
    ev1.record()
    barrier()
    sampler.mark_end()
    launches = _ffi.kernel_launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    import ctypes
    k_ms, k_n = ctypes.c_double(0), ctypes.c_int64(0)
    _ffi.check(_ffi.lib().cdr_prof_read(0, ctypes.byref(k_ms), ctypes.byref(k_n)))
    _ffi.lib().cdr_prof_enable(0)
    # too few clock samples (short region, e.g. 8 GPUs): every rank keeps stepping, untimed, while rank 0 samples
    # (a step COUNT agreed by all ranks: every rank must issue the same sequence of exchanges)
    n_ext = min(20000, int(0.6 / max(ms / 1e3 / max(args.steps, 1), 1e-5)) + 1) if (rank == 0 and sampler.samples_in_region() < 3 and sampler.proc) else 0
    need_more = torch.tensor([n_ext], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.broadcast(need_more, 0)
    if int(need_more.item()):
        for _ in range(int(need_more.item())):
            searcher.search(q_dev[total_steps - 1], TOPK)
        barrier()
        if rank == 0:
            sampler.extended = True
            sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms, float(launches), k_ms.value], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, launches_total, k_ms_max = float(tmax[0]), int(tsum[1]), float(tmax[2])
    else:
        launches_total, k_ms_max = int(launches), k_ms.value
    value = args.steps * Q / (ms / 1e3)

    # informational: the same steps with shared reads (groups of 16 queries -- 3 for short tails -- score every
    # streamed tile; identical results)
    shared = None
    if Q >= 2:
        for s_ in range(min(args.warmup, 3)):
            searcher.search(q_dev[s_], TOPK, shared=True)
        barrier()
        sa, sb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sa.record()
        for s_ in range(args.warmup, total_steps):
            got = searcher.search(q_dev[s_], TOPK, shared=True)
        sb.record()
        barrier()
        sms = torch.tensor([sa.elapsed_time(sb)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(sms, op=dist.ReduceOp.MAX)
        assert torch.equal(got[0], last[0]) and torch.equal(got[1].view(torch.int64), last[1].view(torch.int64)), \
            "shared-read scan differs from one-scan-per-query"
        shared = {"queries_per_pass": 16 if (Q >= 10 and TOPK <= 56) else 3, "value": args.steps * Q / (float(sms[0]) / 1e3), "unit": UNIT,
                  "ms_per_step": float(sms[0]) / args.steps,
                  "note": "cdr_search_exact_f32_shared: not the contract line (configs[1] is one scan per query)"}

    # single-query latency (device-timed, one query per call)
    lat = []
    for i in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); searcher.search(q_dev[args.warmup + i % args.steps][i % Q: i % Q + 1], TOPK); b.record()
        torch.cuda.synchronize()
        lat.append(a.elapsed_time(b))
    lat.sort()

    # ---- timed: end to end with host buffers
    e2e = None
    if not args.no_e2e:
        def e2e_step(s):
            if world == 1:
                return store.search_exact(q_host[s], TOPK)           # cdr_search_exact_f32_host
            qd = torch.from_numpy(q_host[s]).cuda(non_blocking=False)
            ids, sc, n = searcher.search(qd, TOPK)
            return ids.cpu().numpy(), sc.cpu().numpy(), n.cpu().numpy()
        for s in range(min(args.warmup, 3)):
            e2e_step(s)
        barrier()
        t0 = time.perf_counter()
        for s in range(args.warmup, total_steps):
            out = e2e_step(s)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": args.steps * Q / float(tt[0]), "unit": UNIT, "h2d_bytes_per_step": Q * DIM * 4,
               "d2h_bytes_per_step": Q * TOPK * 16 + Q * 4}
        # the host-buffer path returns the same bits as the device path
        assert np.array_equal(out[0], last[0].cpu().numpy()), "e2e ids differ from device-path ids"

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
        launches_k1 = int(k_n.value)
        bytes_per_launch = float(count) * DIM * 4 * Q          # one launch scans the shard for Q queries
        achieved = bytes_per_launch / (k_ms_max / max(launches_k1, 1) / 1e3) / 1e9 if k_ms_max > 0 else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(tpath) and world == 1:
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_row", 0) * count * Q if tj.get("dram_bytes_per_row") else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, world), sharding=f"rows/{world}" if world > 1 else "none",
                           exchange=searcher.transport,
                           l2="inputs larger than L2 (shard bytes >> 126 MB), distinct queries every step",
                           single_query_latency_ms_p50=lat[len(lat) // 2], single_query_latency_ms_min=lat[0]),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_total,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "kernel": "exact_scan_kernel<8,2,2>", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_launch,
                         "avg_launch_ms": k_ms_max / max(launches_k1, 1), "launches_timed": launches_k1},
        }
        line["exact_batch_shared_reads"] = shared
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.rows, args.cpu_sample_rows, args.cpu_sample_queries, 0)
            one = cpu_baseline(args.rows, min(args.cpu_sample_rows, 50_000), 4, 1, target_seconds=4.0)
            line["cpu_baseline"]["single_thread_value"] = one["value"]
            # the reference's other dense path (mode "ann": HNSW, ef_search = 80), restated, in the same run
            line["cpu_baseline_hnsw"] = cpu_baseline_hnsw(args.hnsw_rows, 0, target_seconds=5.0)
        emit(line)
    if world > 1:
        dist.barrier()
    searcher.close()
    store.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
