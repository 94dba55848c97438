"""Host profile of one-at-a-time retrieve_ids (1M chunks): filtered (10 calls = 2 000 rows) and unfiltered."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
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
texts = [f"status of TK-{i % 500} and TK-{(i * 13) % 900}" for i in range(400)]
for name, f in (("filtered", filt), ("unfiltered", None)):
    for t in texts[:20]:
        retrieve.retrieve_ids(eng, t, f)
    t0 = time.perf_counter()
    for t in texts:
        retrieve.retrieve_ids(eng, t, f)
    dt = time.perf_counter() - t0
    print(f"{name}: {len(texts) / dt:.0f} requests/s one at a time ({dt / len(texts) * 1e3:.3f} ms per request)")
    pr = cProfile.Profile(); pr.enable()
    for t in texts:
        retrieve.retrieve_ids(eng, t, f)
    pr.disable()
    sio = io.StringIO(); pstats.Stats(pr, stream=sio).sort_stats("tottime").print_stats(28); print(sio.getvalue()[:6500])
