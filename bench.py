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

The default line also carries `sub_records`, one per other GPU config of BASELINE.json, each with its own
clocks, roofline, e2e and in-run parity:
  sub_records.batch_bf16  configs[2] on 1 GPU (10M x 1024 bf16, 1024 queries per step on the tcgen05 lane);
                          under torchrun configs[4]: the fixed 100M-row corpus row-sharded over the N GPUs
                          (fp32 rows resident too while 6 KB/row fits, i.e. at N = 8; bf16-only at N = 2, 4)
  sub_records.hybrid      configs[3] + configs[0] (N = 1 only): hybrid /retrieve through the facade
`parity` objects: every record checks the results of its last timed step against oracles that share no code
with the engine -- a plain PyTorch fp32 matmul (TF32 off) + fp64 re-score over the regenerated (or stored bf16)
rows, merged across ranks with torch; the C oracle on a row window; at N > 1 the unsharded scan on rank 0 --
and the run exits non-zero WITHOUT a bench line if one of them fails.  `exact_batch_shared_reads` (the same
steps with queries sharing every streamed tile) and the CPU baselines (exact scan + HNSW, restated) ride along.
Standalone lines: --workload batch_bf16 | hybrid | ingest.

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
    ap.add_argument("--reference-rows", type=int, default=0,
                    help="--impl reference: rows the CPU arm scans (default 0 = the WHOLE corpus of the config, no scaling; "
                         "a smaller value scans a sample and scales the rate by rows, and the line says so)")
    ap.add_argument("--batch-queries", type=int, default=1024, help="batch_bf16: queries per batch (configs[2]: 1024)")
    ap.add_argument("--hnsw-rows", type=int, default=0, help="rows of the CPU HNSW baseline's sample (default 0: not run -- "
                                                             "on the iid synthetic corpus a graph index is not a comparable baseline, BASELINE.md 4)")
    ap.add_argument("--batch-rows", type=int, default=0, help="batch_bf16: corpus rows (default 10M on 1 GPU, 100M on N > 1)")
    ap.add_argument("--no-sub-records", action="store_true", help="default line only: skip the configs[2..4] sub-records")
    ap.add_argument("--no-parity", action="store_true", help="A/B timing aids only (e.g. CADENCE_K2_DRYRUN, wrong results by design): "
                                                             "skip the in-run parity checks; the line says so")
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
    # the SAME dict in both arms (the driver compares them): everything that describes this run rather than the workload
    # -- sharding, transport, latencies -- lives under the line's "run" key
    return {"workload": f"{label}: {args.rows} x {DIM} fp32 corpus, single-query exact cosine scan + "
                        f"top-k={TOPK}; step = {args.queries_per_step} distinct queries, one scan of the corpus per query",
            "rows": args.rows, "dim": DIM, "k": TOPK, "queries_per_step": args.queries_per_step,
            "l2": "inputs larger than L2 (corpus bytes >> 126 MB), distinct queries every step"}


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
    rows = args.rows if args.reference_rows <= 0 else min(args.reference_rows, args.rows)
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
    scaled = "" if rows == args.rows else ", rate scaled by rows"
    sample = (f"{q_per_step} queries/step x {rows} of {args.rows} rows{scaled}; pgvector 0.8.1 "
              f"exact-scan loop restated in C (oracle/pgvector_restated.c), {cores} OpenMP threads; no Postgres (no TOAST / "
              f"tuple / executor overhead): optimistic for the reference")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


# ----------------------------------------------------------------------------- B200 arm: shared plumbing
class Ctx:
    """One rank's view of the job: torch.distributed over NCCL when launched under torchrun."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback "
                             "(use --impl reference for the CPU baseline)")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.local_rank}"))
        # the independent oracle below is a plain fp32 matmul: no TF32 anywhere in this process
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op="max"):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def all_ok(self, ok: bool) -> bool:
        """AND of a per-rank verdict over all ranks: a check only rank 0 can make (it holds the oracle's answer) must
        fail on EVERY rank, or the others would walk on into the next collective."""
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(int(t.item()))

    def agree(self, n: int) -> int:
        """rank 0's integer on every rank (step counts every rank must share)."""
        t = self.torch.tensor([n], dtype=self.torch.int64, device="cuda")
        if self.world > 1:
            self.dist.broadcast(t, 0)
        return int(t.item())

    def close(self):
        if self.world > 1 and self.dist.is_initialized():
            self.dist.barrier()
            self.dist.destroy_process_group()


def extend_for_clocks(ctx, sampler, ms, steps, step_fn):
    """Too few clock samples (short region, e.g. 8 GPUs): every rank keeps stepping, untimed, while rank 0 samples
    (a step COUNT agreed by all ranks: every rank must issue the same sequence of exchanges)."""
    n_ext = 0
    if ctx.rank == 0 and sampler.proc and sampler.samples_in_region() < 3:
        n_ext = min(20000, int(0.6 / max(ms / 1e3 / max(steps, 1), 1e-5)) + 1)
    n_ext = ctx.agree(n_ext)
    if n_ext:
        for _ in range(n_ext):
            step_fn()
        ctx.barrier()
        if ctx.rank == 0:
            sampler.extended = True
            sampler.mark_end()


# ----------------------------------------------------------------------------- in-run parity: independent oracles
class ParityError(AssertionError):
    """An in-run parity check failed: the run exits non-zero and prints no bench line."""


def oracle_topk_torch(ctx, chunks, q, k, id_base=1):
    """Top-k by cosine with plain PyTorch, sharing no code with the engine or with oracle/ (SURVEY 8(d) C3/C5):
    per chunk of rows an fp32 matmul (TF32 off) keeps the best k+14 candidates and their vectors; the survivors
    are re-scored in fp64 with pgvector's formula and ordered (score desc, id asc).  `chunks` yields
    (global_first_row, x fp32 [m, D] CUDA tensor) over THIS rank's rows; with several ranks the per-rank lists
    are all-gathered and merged with torch as well.  Returns (ids [nq,k] int64, scores [nq,k] float64)."""
    torch = ctx.torch
    kc = k + 14
    nq = q.shape[0]
    qn = q / q.norm(dim=1, keepdim=True)
    best_s = torch.full((nq, 0), float("-inf"), device=q.device)
    best_r = torch.zeros((nq, 0), dtype=torch.int64, device=q.device)
    best_v = torch.zeros((nq, 0, q.shape[1]), device=q.device)
    for first, x in chunks:
        s = (qn @ x.T) / x.norm(dim=1)[None, :]
        s = torch.nan_to_num(s, nan=float("-inf"))
        cs, ci = torch.topk(s, min(kc, x.shape[0]), dim=1)
        best_s = torch.cat([best_s, cs], dim=1)
        best_r = torch.cat([best_r, ci + first], dim=1)
        best_v = torch.cat([best_v, x[ci]], dim=1)
        if best_s.shape[1] > kc:
            best_s, sel = torch.topk(best_s, kc, dim=1)
            best_r = torch.gather(best_r, 1, sel)
            best_v = torch.gather(best_v, 1, sel[:, :, None].expand(-1, -1, best_v.shape[2]))
        del s, x
    a, b = q.double(), best_v.double()
    ab = (a[:, None, :] * b).sum(dim=2)
    sim = ab / torch.sqrt((a * a).sum(dim=1)[:, None] * (b * b).sum(dim=2))
    sim = sim.clamp(-1.0, 1.0)
    score = 1.0 - (1.0 - sim)
    ids = best_r + id_base
    if ctx.world > 1:
        pad = kc - score.shape[1]                      # a short shard: pad so every rank gathers the same shape
        if pad > 0:
            score = torch.cat([score, torch.full((nq, pad), float("-inf"), dtype=score.dtype, device=q.device)], dim=1)
            ids = torch.cat([ids, torch.full((nq, pad), 2 ** 62, dtype=ids.dtype, device=q.device)], dim=1)
        g_s = [torch.empty_like(score) for _ in range(ctx.world)]
        g_i = [torch.empty_like(ids) for _ in range(ctx.world)]
        ctx.dist.all_gather(g_s, score.contiguous())
        ctx.dist.all_gather(g_i, ids.contiguous())
        score, ids = torch.cat(g_s, dim=1), torch.cat(g_i, dim=1)
    o = torch.argsort(ids, dim=1, stable=True)
    score, ids = torch.gather(score, 1, o), torch.gather(ids, 1, o)
    o = torch.argsort(score, dim=1, descending=True, stable=True)
    return torch.gather(ids, 1, o)[:, :k].contiguous(), torch.gather(score, 1, o)[:, :k].contiguous()


def synth_chunks(seed, first, count, device, chunk=1 << 20):
    """This rank's rows regenerated from the counter-based generator (fp32), chunk by chunk."""
    from cadence_rag_b200.store import synth_rows_device
    for off in range(0, count, chunk):
        m = min(chunk, count - off)
        yield first + off, synth_rows_device(seed, first + off, m, DIM, device=device)


def stored_bf16_chunks(store, first, count, chunk=1 << 20):
    """This rank's rows AS STORED in bf16 (bf16-only residency: the corpus is the bf16 values), widened to fp32."""
    for off in range(0, count, chunk):
        m = min(chunk, count - off)
        yield first + off, store.read_rows_device(off, m, "bf16").float()


def compare_lists(torch, got_ids, got_sc, want_ids, want_sc, k):
    """recall@k, the share of identical positions, and the largest relative score error on matching positions."""
    g, w = got_ids.cpu().numpy(), want_ids.cpu().numpy()
    import numpy as np
    recall = float(np.mean([len(set(g[i].tolist()) & set(w[i].tolist())) / k for i in range(g.shape[0])]))
    same = g == w
    gs, ws = got_sc.cpu().numpy(), want_sc.cpu().numpy()
    rel = np.abs(gs - ws) / np.maximum(np.abs(ws), 1e-300)
    return recall, float(same.mean()), float(rel[same].max()) if same.any() else None


_X_WINDOW = {}


def window_rows_host(rows):
    """Global rows [0, rows) of the synthetic corpus from the C generator (cached: the CPU baseline uses them too)."""
    from oracle import cpu_oracle as orc
    if rows not in _X_WINDOW:
        _X_WINDOW.clear()
        _X_WINDOW[rows] = orc.synth_rows(20260209, 0, rows)
    return _X_WINDOW[rows]


def c_oracle_window_check(ctx, searcher, first, count, q_dev, k, mode, window, stored_bf16=False, n_queries=4):
    """The C oracle (oracle/pgvector_restated.c, fp64 variant = ground-truth order) on global rows [0, window) against
    the engine restricted to the same rows by an allow bitmap -- the (sharded) lane under test, filter path included.
    Every rank takes part in the search; rank 0 compares.  Returns a dict for the bench line; raises ParityError."""
    import numpy as np
    torch = ctx.torch
    from oracle import cpu_oracle as orc
    hi = max(0, min(count, window - first))                             # local rows [0, hi) lie inside the window
    bits = np.zeros(((count + 31) // 32) * 32, dtype=bool)
    bits[:hi] = True
    allow = torch.from_numpy(orc.rows_to_bitmap(bits).view(np.int32)).cuda()
    q = q_dev[:n_queries].contiguous()
    ids, sc, n = searcher.search(q, k, allow=allow, mode=mode)
    ctx.barrier()
    if ctx.rank != 0:
        return None
    qh = q.cpu().numpy()
    if stored_bf16:
        m = min(window, count)
        xb = searcher.store.read_rows(0, m, ("bf16",))["bf16"]
        want = [orc.exact_scan_bf16rows(qh[i], xb, k) for i in range(n_queries)]
        window = m
    else:
        x = window_rows_host(window)
        want = [orc.exact_scan(qh[i], x, k, variant=orc.VARIANT_F64) for i in range(n_queries)]
    g_ids, g_sc, g_n = ids.cpu().numpy(), sc.cpu().numpy(), n.cpu().numpy()
    same, recall, worst = 0, 0.0, 0.0
    for i, (w_ids, w_sc) in enumerate(want):
        m = len(w_ids)
        if int(g_n[i]) != m:            # reported through the metrics (rank 0 alone must not raise: see Ctx.all_ok)
            return {"oracle": "C restatement", "queries": n_queries, "recall_at_k": 0.0, "identical_positions": 0.0,
                    "max_rel_score_err": float("inf"),
                    "mismatch": f"{mode}: window query {i}: {int(g_n[i])} results, the C oracle has {m}"}
        recall += len(set(g_ids[i, :m].tolist()) & set(w_ids.tolist())) / max(m, 1)
        eq = g_ids[i, :m] == w_ids
        same += int(eq.sum())
        if eq.any():
            worst = max(worst, float(np.max(np.abs(g_sc[i, :m][eq] - w_sc[eq]) / np.maximum(np.abs(w_sc[eq]), 1e-300))))
    recall /= n_queries
    ident = same / float(sum(len(w[0]) for w in want))
    return {"oracle": "C restatement (oracle/pgvector_restated.c, fp64 accumulate" + (", bf16-valued rows as stored)" if stored_bf16 else ")"),
            "rows": f"global rows [0, {window}) selected by an allow bitmap", "queries": n_queries,
            "recall_at_k": recall, "identical_positions": ident, "max_rel_score_err": worst}


# ----------------------------------------------------------------------------- configs[1]: exact fp32 scan (the contract line)
def record_exact_f32(args, ctx, keep_store=False):
    """BASELINE configs[1]: returns (line on rank 0 / None elsewhere, store or None)."""
    import ctypes
    import numpy as np
    torch, dist = ctx.torch, ctx.dist
    from cadence_rag_b200 import _ffi
    from cadence_rag_b200.dist import ShardedSearcher, shard_range
    from cadence_rag_b200.store import DenseStore, SYNTH_CORPUS_SEED, SYNTH_QUERY_SEED, synth_rows_device
    world, rank, local_rank = ctx.world, ctx.rank, ctx.local_rank
    _ffi.require_device()

    Q = args.queries_per_step
    first, count = shard_range(args.rows, rank, world)
    store = DenseStore("chunks", max(count, 1), dim=DIM, device=local_rank, fp32=True, bf16=keep_store)
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
    barrier = ctx.barrier

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
    k_ms, k_n = ctypes.c_double(0), ctypes.c_int64(0)
    _ffi.check(_ffi.lib().cdr_prof_read(0, ctypes.byref(k_ms), ctypes.byref(k_n)))
    _ffi.lib().cdr_prof_enable(0)
    extend_for_clocks(ctx, sampler, ms, args.steps, lambda: searcher.search(q_dev[total_steps - 1], TOPK))
    clocks = sampler.stop() if rank == 0 else None

    ms, k_ms_max = ctx.reduce([ms, k_ms.value], "max")
    launches_total = int(ctx.reduce([float(launches)], "sum")[0])
    value = args.steps * Q / (ms / 1e3)

    # ---- in-run parity (the run fails if any of it fails)
    parity = {}
    nchk = min(32, Q)
    q_chk = q_dev[total_steps - 1][:nchk].contiguous()
    o_ids, o_sc = oracle_topk_torch(ctx, synth_chunks(SYNTH_CORPUS_SEED, first, count, local_rank), q_chk, TOPK)
    recall, ident, rel = compare_lists(torch, last[0][:nchk], last[1][:nchk], o_ids, o_sc, TOPK)
    parity["torch_fp32_matmul_fp64_rescore"] = {
        "oracle": "plain PyTorch: fp32 matmul (TF32 off) top-64 candidates per 1M-row chunk of the regenerated corpus, fp64 "
                  "re-score, order (score desc, id asc); per-rank lists merged with torch",
        "queries": nchk, "rows": args.rows, "recall_at_50": recall, "identical_positions": ident, "max_rel_score_err": rel}
    if ident != 1.0 or (rel is not None and rel > 1e-9):
        raise ParityError(f"exact fp32 lane differs from the torch oracle: identical positions {ident}, recall {recall}, "
                          f"max rel score err {rel}")
    if not args.no_cpu_baseline:
        win = c_oracle_window_check(ctx, searcher, first, count, q_chk, TOPK, "exact", min(args.cpu_sample_rows, args.rows))
        parity["c_oracle_window"] = win
        if not ctx.all_ok(rank != 0 or (win["identical_positions"] == 1.0 and win["max_rel_score_err"] <= 1e-9)):
            raise ParityError(f"exact fp32 lane differs from the C oracle on the row window: {win}")
    if world > 1:
        # multi-GPU merge parity: the same queries over the WHOLE corpus on rank 0 alone must give the same bits
        verdict = torch.ones(1, dtype=torch.int32, device="cuda")
        if rank == 0:
            whole = DenseStore("chunks", args.rows, dim=DIM, device=local_rank, fp32=True, bf16=False)
            whole.append_synthetic(args.rows)
            whole.finalize()
            w_ids, w_sc, w_n = whole.search_exact(q_dev[total_steps - 1], TOPK)
            ok = (torch.equal(w_ids, last[0]) and torch.equal(w_sc.view(torch.int64), last[1].view(torch.int64))
                  and torch.equal(w_n, last[2]))
            verdict[0] = 1 if ok else 0
            whole.close()
        dist.broadcast(verdict, 0)
        if int(verdict.item()) != 1:
            raise ParityError(f"sharded merge over {world} ranks differs from the unsharded scan of the same corpus")
        parity["sharded_vs_unsharded_same_bits"] = {"queries": Q, "ranks": world, "identical": True,
                                                    "note": "rank 0 scans the whole corpus alone: ids, score bits and counts equal"}

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
        sms = ctx.reduce([sa.elapsed_time(sb)], "max")[0]
        if not (torch.equal(got[0], last[0]) and torch.equal(got[1].view(torch.int64), last[1].view(torch.int64))):
            raise ParityError("shared-read scan differs from one-scan-per-query")
        shared = {"queries_per_pass": 16 if (Q >= 10 and TOPK <= 56) else 3, "value": args.steps * Q / (sms / 1e3), "unit": UNIT,
                  "ms_per_step": sms / args.steps,
                  "note": "cdr_search_exact_f32_shared: not the contract line (configs[1] is one scan per query)"}

    # single-query latency (device-timed, one query per call)
    lat = []
    for i in range(40):
        q1 = q_dev[args.warmup + i % args.steps][i % Q: i % Q + 1]      # the request's vector, resident before the call
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); searcher.search(q1, TOPK); b.record()
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
        dt = ctx.reduce([time.perf_counter() - t0], "max")[0]
        e2e = {"value": args.steps * Q / dt, "unit": UNIT, "h2d_bytes_per_step": Q * DIM * 4,
               "d2h_bytes_per_step": Q * TOPK * 16 + Q * 4}
        # the host-buffer path returns the same bits as the device path
        if not np.array_equal(out[0], last[0].cpu().numpy()):
            raise ParityError("e2e ids differ from device-path ids")

    line = None
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
        launches_k1 = int(k_n.value)
        # a step scans the shard once per query; its K1 time is one launch (N = 1) or one launch per 16-query chunk
        # (the pipelined sharded step): algorithmic bytes of all timed steps / total K1 time of those steps
        per_step = max(1, launches_k1 // max(args.steps, 1))
        bytes_per_launch = float(count) * DIM * 4 * Q / per_step
        achieved = float(count) * DIM * 4 * Q * args.steps / (k_ms_max / 1e3) / 1e9 if k_ms_max > 0 else None
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("dram_bytes_per_row"):
                traffic = tj["dram_bytes_per_row"] * count * Q
                traffic_src = ("committed ncu --set full capture of this kernel (profiles/k1_traffic.json: DRAM bytes per row), scaled to this launch"
                               + ("" if world == 1 else "; per GPU: the same kernel over this rank's shard (the capture is a 1-GPU one)"))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "run": {"sharding": f"rows/{world}" if world > 1 else "none", "exchange": searcher.transport,
                    "single_query_latency_ms_p50": lat[len(lat) // 2], "single_query_latency_ms_min": lat[0]},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_total,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "exact_scan_kernel<8,2,2>", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_launch,
                         "avg_launch_ms": k_ms_max / max(launches_k1, 1), "launches_timed": launches_k1,
                         "launches_per_step": per_step, "k1_ms_per_step": k_ms_max / max(args.steps, 1)},
            "parity": parity,
        }
        line["exact_batch_shared_reads"] = shared
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.rows, args.cpu_sample_rows, args.cpu_sample_queries, 0)
            one = cpu_baseline(args.rows, min(args.cpu_sample_rows, 50_000), 4, 1, target_seconds=4.0)
            line["cpu_baseline"]["single_thread_value"] = one["value"]
            if args.hnsw_rows > 0:
                # the reference's other dense path (mode "ann": HNSW, ef_search = 80), restated, in the same run
                line["cpu_baseline_hnsw"] = cpu_baseline_hnsw(args.hnsw_rows, 0, target_seconds=5.0)
            else:
                line["cpu_baseline_hnsw"] = {
                    "not_comparable": "HNSW (m=16, ef_construction=64, ef_search=80, restated) on the iid 1024-d synthetic corpus reaches "
                                      "recall@50 0.31 at 10 000 rows, 0.096 at 50 000 and 0.026 at 200 000 (profiles/r01/hnsw_cpu_sweep.json), "
                                      "and a 1M-row build takes hours on host cores: see BASELINE.md 4; --hnsw-rows N times it on an N-row sample"}
    barrier()
    searcher.close()
    if keep_store and world == 1:
        return line, store
    store.close()
    return line, None


# ----------------------------------------------------------------------------- configs[2] / configs[4]: batched bf16 lane
def batch_bf16_rows(args, world):
    """Corpus of the batched lane: configs[2] = 10M rows on one GPU; configs[4] = the fixed 100M-row corpus
    row-sharded over 2 / 4 / 8 GPUs (strong scaling on 100M rows)."""
    if args.batch_rows > 0:
        return args.batch_rows
    return 10_000_000 if world == 1 else 100_000_000


def _batch_parity(args, ctx, parity, store, searcher, first, count, per, rows, bf16_only, q_chk, got_ids, got_sc, nchk):
    """In-run parity of the batched lane against independent oracles (fills `parity`, raises ParityError)."""
    from cadence_rag_b200.store import SYNTH_CORPUS_SEED
    torch = ctx.torch
    rank, world, local_rank = ctx.rank, ctx.world, ctx.local_rank
    desc = ("plain PyTorch: fp32 matmul (TF32 off) top-64 candidates per 1M-row chunk, fp64 re-score, order (score desc, "
            "id asc); per-rank lists merged with torch; ")
    if bf16_only:
        o_ids, o_sc = oracle_topk_torch(ctx, stored_bf16_chunks(store, first, count), q_chk, TOPK)
        recall, ident, rel = compare_lists(torch, got_ids, got_sc, o_ids, o_sc, TOPK)
        parity["torch_fp32_matmul_over_bf16_valued_rows"] = {
            "oracle": desc + "rows = the stored bf16 values widened to fp32 (bf16-only residency: fp32 query x bf16-valued rows)",
            "queries": nchk, "rows": rows, "recall_at_50": recall, "identical_positions": ident, "max_rel_score_err": rel}
        if recall < 0.999:
            raise ParityError(f"bf16 lane (bf16-only store) recall {recall} < 0.999 against the torch oracle on the stored rows")
    f_ids, f_sc = oracle_topk_torch(ctx, synth_chunks(SYNTH_CORPUS_SEED, first, count, local_rank), q_chk, TOPK)
    recall32, ident32, rel32 = compare_lists(torch, got_ids, got_sc, f_ids, f_sc, TOPK)
    parity["torch_fp32_matmul_over_fp32_rows"] = {
        "oracle": desc + "rows = the fp32 corpus regenerated from the counter-based generator (the north-star truth: "
                         "pgvector exact over the fp32 embeddings)",
        "queries": nchk, "rows": rows, "recall_at_50": recall32, "identical_positions": ident32,
        "max_rel_score_err": None if bf16_only else rel32,
        "note": "bf16-only residency: the engine's re-score sees bf16-valued rows, so recall against the fp32 truth is "
                "bounded by bf16 rounding (reported, not asserted)" if bf16_only else None}
    if not bf16_only and recall32 < 0.999:
        raise ParityError(f"bf16 lane recall {recall32} < 0.999 against the torch fp32 oracle")
    if not args.no_cpu_baseline:
        win = c_oracle_window_check(ctx, searcher, first, count, q_chk, TOPK, "ann", min(args.cpu_sample_rows, per),
                                    stored_bf16=bf16_only)
        parity["c_oracle_window"] = win
        if not ctx.all_ok(rank != 0 or win["recall_at_k"] >= 0.999):
            raise ParityError(f"bf16 lane differs from the C oracle on the row window: {win}")


def record_batch_bf16(args, ctx):
    """BASELINE configs[2] on 1 GPU, configs[4] row-sharded on N GPUs: rows x 1024 bf16 corpus, 1024 queries per
    step on the tcgen05 lane (K2) + exact re-score; with N > 1 the per-rank top-k lists are exchanged and merged.
    roofline: tensor-bound, 2*nq*rows_local*dim FLOP per step per GPU.  Returns the record on rank 0."""
    import ctypes
    import numpy as np
    torch = ctx.torch
    from cadence_rag_b200 import _ffi
    from cadence_rag_b200.dist import ShardedSearcher, shard_range
    from cadence_rag_b200.store import DenseStore, SYNTH_CORPUS_SEED, SYNTH_QUERY_SEED, synth_rows_device
    world, rank, local_rank = ctx.world, ctx.rank, ctx.local_rank
    barrier = ctx.barrier
    rows = batch_bf16_rows(args, world)
    nq = args.batch_queries
    first, count = shard_range(rows, rank, world)
    # fp32 rows stay resident next to the bf16 copy while both fit comfortably (6 KB per row, <= 110 GB per GPU);
    # above that the shard is bf16-only (2 and 4 GPUs on the 100M corpus) and "exact" means exact
    # over the bf16-valued rows
    per = (rows + world - 1) // world
    bf16_only = args.bf16_only or per * DIM * 6 > 110e9
    store = DenseStore("chunks", max(count, 1), dim=DIM, device=local_rank, fp32=not bf16_only, bf16=True)
    store.append_synthetic(count, first_row=first)
    store.finalize()
    torch.cuda.synchronize()
    searcher = ShardedSearcher(store)
    total = args.warmup + args.steps
    q_dev = synth_rows_device(SYNTH_QUERY_SEED, 0, total * nq, DIM, device=local_rank).view(total, nq, DIM)

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
    extend_for_clocks(ctx, sampler, ms, args.steps, lambda: searcher.search(q_dev[total - 1], TOPK, mode="ann"))
    clocks = sampler.stop() if rank == 0 else None
    ms, gemm_ms = ctx.reduce([ms, k_ms.value], "max")
    launches = int(ctx.reduce([float(launches)], "sum")[0])

    # ---- in-run parity against independent oracles, 32 queries of the last batch (the run fails if it fails)
    parity = {}
    nchk = min(32, nq)
    q_chk = q_dev[total - 1][:nchk].contiguous()
    got_ids, got_sc = out[0][:nchk], out[1][:nchk]
    if args.no_parity:
        parity = {"skipped": "--no-parity (timing aid run)"}
    else:
        _batch_parity(args, ctx, parity, store, searcher, first, count, per, rows, bf16_only, q_chk, got_ids, got_sc, nchk)
    recall_exact_lane = None
    if store.has_fp32:
        nr = min(64, nq)
        e_ids, _, _ = searcher.search(q_dev[total - 1][:nr].contiguous(), TOPK, mode="exact")
        got = out[0][:nr].cpu().numpy(); want = e_ids.cpu().numpy()
        recall_exact_lane = float(np.mean([len(set(got[i]) & set(want[i])) / TOPK for i in range(nr)]))

    def q1_latency(mode, bytes_per_row):
        for i in range(3):
            searcher.search(q_dev[0][i % nq:i % nq + 1].contiguous(), TOPK, mode=mode)
        barrier()
        xa, xb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        xa.record()
        for i in range(20):
            got1 = searcher.search(q_dev[1][i % nq:i % nq + 1].contiguous(), TOPK, mode=mode)
        xb.record()
        barrier()
        xt = ctx.reduce([xa.elapsed_time(xb) / 20], "max")[0]
        return {"ms_per_query": xt, "hbm_gbs_per_gpu": float(count) * DIM * bytes_per_row / (xt / 1e3) / 1e9}, got1
    # the exact fp32 lane over the same (sharded) corpus, one query per request: device-timed latency
    exact_q1 = q1_latency("exact", 4)[0] if store.has_fp32 else None
    # mode "ann" for ONE query: the same scan over the bf16 rows (half the bytes), exact re-score
    scan_q1, got1 = q1_latency("scan_bf16", 2)
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
    e2e = None
    if not args.no_e2e:
        e2e_step(0)
        barrier()
        t0 = time.perf_counter()
        for s in range(args.warmup, total):
            eo = e2e_step(s)
        barrier()
        dt = ctx.reduce([time.perf_counter() - t0], "max")[0]
        e2e = {"value": args.steps * nq / dt, "unit": UNIT, "h2d_bytes_per_step": nq * DIM * 4,
               "d2h_bytes_per_step": nq * TOPK * 16 + nq * 4}
        if not args.no_parity and not np.array_equal(eo[0], out[0].cpu().numpy()):
            raise ParityError("batched lane: e2e ids differ from device-path ids")
    line = None
    if rank == 0:
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peaks = json.load(open(pk)) if os.path.exists(pk) else {}
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        burst = peaks.get("bf16_tflops", 1646.0)
        flops_step_gpu = 2.0 * nq * per * DIM               # the largest shard (rank 0)
        # DRAM bytes of all gemm_topk launches of one step, from an ncu capture of this command
        # (profiles/k2_traffic.json: bytes per corpus row), scaled to this shard
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "k2_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("dram_bytes_per_row") and tj.get("queries") == nq:
                traffic = tj["dram_bytes_per_row"] * per
        achieved = flops_step_gpu * args.steps / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        line = {"metric": "queries/sec (top-k=50, 1024-d) batched bf16 tcgen05 lane",
                "value": args.steps * nq / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"BASELINE configs[{2 if world == 1 else 4}]: {rows} x {DIM} bf16 corpus"
                                       f"{' row-sharded over %d GPUs' % world if world > 1 else ''}, batch {nq} queries, "
                                       f"tcgen05 GEMM with fused threshold top-k epilogue + exact re-score, top-k={TOPK}",
                           "rows": rows, "rows_per_gpu": per, "dim": DIM, "k": TOPK, "queries_per_step": nq,
                           "resident": "bf16 only" if bf16_only else "fp32 + bf16",
                           "l2": "inputs larger than L2", "recall_at_50_vs_exact_fp32_lane": recall_exact_lane,
                           "exchange": searcher.transport, "exact_fp32_lane_single_query": exact_q1,
                           "ann_bf16_scan_single_query": scan_q1},
                "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": achieved / peak if achieved else None, "traffic": traffic,
                             "frac_of_burst_peak": achieved / burst if achieved else None,
                             "algorithmic_flops_per_step": flops_step_gpu,
                             "corpus_stream_gbs": (float(per) * DIM * 2 * args.steps / (gemm_ms / 1e3) / 1e9) if gemm_ms > 0 else None,
                             "algorithmic_dram_bytes_per_step": float(per) * DIM * 2,
                             "kernel": "gemm_topk_kernel", "per": "GPU (max over ranks)",
                             "peak_source": "measured bf16_tflops_sustained (kernel timed inside a long step)",
                             "peak_burst": burst, "gemm_ms_per_step": gemm_ms / args.steps,
                             "launches_timed": int(k_n.value), "segment_launch_ms_last_step": last_step_launch_ms,
                             "gemm_ms_by_step": [round(float(per_launch[i * segs:(i + 1) * segs].sum()), 3)
                                                 for i in range(args.steps)]},
                "parity": parity}
    barrier()
    searcher.close()
    store.close()
    torch.cuda.empty_cache()
    return line


def run_hybrid(args, store=None, lean=False):
    """Secondary line (BASELINE configs[3] + configs[0]): hybrid /retrieve over 1M chunks through the
    facade -- synthetic embedder, K6 filter bitmap + count, planner, K1 exact scan, host tech_tokens
    lane, K5 RRF -- for (a) a 10-call filter (2 000 candidate rows => mode "exact", the C1 shape) and
    (b) no filter.  Fused ranks are checked bit-exact against the restated pipeline on 8 queries.
    lean=True (the sub-record of the default line): parity, one-at-a-time and 64-per-call rates, the C1 shape;
    the thread / micro-batcher client experiments only run under --workload hybrid.  Returns the record."""
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
    own_store = store is None
    if own_store:
        store = DenseStore("chunks", rows, dim=DIM, device=0, fp32=True, bf16=True)
        store.append_synthetic(rows)
        store.finalize()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = _ffi.kernel_launch_count()
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
        if [(r["chunk_id"], sorted(h), s_) for r, h, s_ in want] != [tuple(t) for t in got["debug"]["fused"]["chunks"]]:
            raise ParityError(f"hybrid: fused ranks of query {qi} differ from the restated pipeline (oracle dense + RRF port)")
        if got["debug"]["dense"]["modes"]["chunks"] != "exact" or got["debug"]["dense"]["candidate_rows"]["chunks"] != 2000:
            raise ParityError(f"hybrid: planner / COUNT(*) of query {qi}: {got['debug']['dense']}")
    sampler.mark_begin()
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
    sampler.mark_end()
    launches = _ffi.kernel_launch_count() - launches0
    clocks = sampler.stop()
    import threading
    for name, f in (() if lean else (("filtered_10_calls_2000_rows", filt), ("unfiltered", None))):
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
    for name, f_of in (() if lean else (("filtered_10_calls_2000_rows", lambda t: filt), ("unfiltered", lambda t: None),
                       ("distinct_filters_per_client", lambda t: RetrieveFilters(call_ids=list(range(10 * t, 10 * t + 10)))))):
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
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": out["unfiltered"]["queries_per_s"], "unit": UNIT,
                    "note": "every rate of this record is end to end: request text in, ids out, host buffers, one sync per call",
                    "h2d_bytes_per_step": DIM * 4 + 256, "d2h_bytes_per_step": 3 * TOPK * 24},
            "roofline": {"bound": "hbm", "kernel": "exact_scan_kernel (unfiltered requests: one scan of the bf16 or fp32 rows each; "
                                                   "filtered: gather launch over 2 000 rows, latency-bound)",
                         "achieved": rows * DIM * 2 * out["unfiltered"]["queries_per_s"] / 1e9, "unit": "GB/s",
                         "note": "whole-request rate x the bf16 scan's bytes: a lower bound on the kernel's own rate (host time included)"},
            "hybrid": out,
            "parity": {"fused_ranks_bit_exact_queries": 8,
                       "oracle": "C oracle dense ids (fp64 variant) over the 2 000 candidate rows + restated tech lane SQL + the RRF port "
                                 "(itself pinned to the reference's _rrf_merge goldens): ids, lane sets and fp64 scores equal"},
            "c1_exact_scan_2000_rows": {"gpu_device_ms": gpu_ms, "gpu_host_buffers_ms": host_ms, "cpu_ms": cpu}}
    embeddings.set_embedder(None)
    for dev_ix in eng.device_tech_indexes.values():
        dev_ix.close()
    if own_store:
        store.close()
    small.close()
    return line


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
    dev_index.close(); store.close(); arts.close()
    return line


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
    if args.workload == "hybrid":
        emit(run_hybrid(args))
        return 0
    if args.workload == "ingest":
        emit(run_ingest(args))
        return 0
    ctx = Ctx()
    try:
        if args.workload == "batch_bf16":
            line = record_batch_bf16(args, ctx)
            if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
                # the reference's two CPU paths for this config, timed on this box's host cores in the same run
                line["cpu_baseline"] = cpu_baseline(batch_bf16_rows(args, 1), args.cpu_sample_rows, args.cpu_sample_queries, 0)
                if args.hnsw_rows > 0:
                    line["cpu_baseline_hnsw"] = cpu_baseline_hnsw(args.hnsw_rows, 0)
        else:
            # the contract line (configs[1]) + one sub-record per other GPU config of BASELINE.json, each with its own
            # clocks, roofline, e2e and in-run parity against an independent oracle
            subs = not args.no_sub_records and args.rows == N_ROWS
            line, store = record_exact_f32(args, ctx, keep_store=subs and ctx.world == 1)
            sub = {}

            def guarded(name, fn):
                """A sub-record whose parity check fails (or that cannot run) carries no number: the headline line is
                still printed, with the failure spelled out where the record would be.  (A failure of the headline
                record's own parity is fatal: no line at all.)"""
                try:
                    return fn()
                except ParityError as exc:
                    return {"value": None, "invalid": True, "parity_failed": str(exc)}
                except Exception as exc:   # noqa: BLE001 - e.g. out of device memory on a smaller part
                    if ctx.world > 1:
                        raise               # ranks must not diverge: a collective would hang
                    return {"value": None, "invalid": True, "error": repr(exc)}
            if subs:
                if ctx.world == 1:
                    sub["hybrid"] = guarded("hybrid", lambda: run_hybrid(args, store=store, lean=True))   # configs[3] (+ configs[0])
                    store.close()
                    ctx.torch.cuda.empty_cache()
                rec = guarded("batch_bf16", lambda: record_batch_bf16(args, ctx))  # configs[2] / configs[4]
                if ctx.rank == 0:
                    sub["batch_bf16"] = rec
                    if ctx.world > 1:
                        sub["hybrid"] = {"skipped": "configs[3] is a single-GPU config: see the N=1 line"}
            if ctx.rank == 0:
                line["sub_records"] = sub
        if ctx.rank == 0:
            emit(line)
    finally:
        ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
