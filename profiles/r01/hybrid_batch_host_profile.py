"""Host profile of retrieve_ids_batch (64 requests per call, 1M chunks): where the Python share of a batched
request goes.  Prints requests/s without the profiler, then the cProfile table."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from cadence_rag_b200 import embeddings, retrieve
from cadence_rag_b200.config import settings
from cadence_rag_b200.lexical import TechTokenIndex
from cadence_rag_b200.retrieve import DenseEngine, RetrieveFilters
from cadence_rag_b200.store import DenseStore, SYNTH_CORPUS_SEED, SYNTH_QUERY_SEED

rows, DIM = 1_000_000, 1024
store = DenseStore("chunks", rows, dim=DIM, device=0, fp32=True, bf16=False)
store.append_synthetic(rows); store.finalize()
rng = np.random.default_rng(SYNTH_CORPUS_SEED)
vocab = 10_000
ntok = rng.integers(0, 4, size=rows)
tok = np.minimum(rng.zipf(1.1, size=(rows, 3)) - 1, vocab - 1)
mask = (np.arange(3)[None, :] < ntok[:, None]).reshape(-1)
flat_rows, flat_tok = np.repeat(np.arange(rows), 3)[mask], tok.reshape(-1)[mask]
order = np.lexsort((flat_rows, flat_tok)); flat_rows, flat_tok = flat_rows[order], flat_tok[order]
starts = np.searchsorted(flat_tok, np.arange(vocab + 1))
index = TechTokenIndex()
for t in range(vocab):
    if starts[t + 1] > starts[t]:
        index.add_postings(f"TK-{t}", np.unique(flat_rows[starts[t]:starts[t + 1]]))
eng = DenseEngine(); eng.register(store, index)
embeddings.set_embedder(embeddings.SyntheticEmbedder(seed=SYNTH_QUERY_SEED, dim=DIM))
settings.embeddings_dim = DIM
filt = RetrieveFilters(call_ids=list(range(10)))
B = 64
batches = [[f"status of TK-{(b * B + i) % 500} and TK-{((b * B + i) * 13) % 900}" for i in range(B)] for b in range(8)]
for name, f in (("unfiltered", None), ("filtered", filt)):
    for b in batches[:2]:
        retrieve.retrieve_ids_batch(eng, b, f)
    t0 = time.perf_counter()
    for b in batches:
        retrieve.retrieve_ids_batch(eng, b, f)
    dt = time.perf_counter() - t0
    print(f"{name}: {len(batches) * B / dt:.0f} requests/s through retrieve_ids_batch ({dt / len(batches) * 1e3:.2f} ms per 64-request call)")
    pr = cProfile.Profile(); pr.enable()
    for b in batches:
        retrieve.retrieve_ids_batch(eng, b, f)
    pr.disable()
    sio = io.StringIO(); pstats.Stats(pr, stream=sio).sort_stats("tottime").print_stats(18); print(sio.getvalue()[:4500])
