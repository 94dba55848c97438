"""Concurrent client threads (own CUDA stream each, one request at a time per thread) through retrieve_ids over 1M chunks.
CADENCE_SYNC was an experimental switch for how a request thread waits for its stream (see concurrent_clients_probe.json); the shipped library always spins."""
import json, os, sys, threading, time
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from cadence_rag_b200 import embeddings, retrieve
from cadence_rag_b200.config import settings
from cadence_rag_b200.lexical import TechTokenIndex
from cadence_rag_b200.retrieve import DenseEngine, RetrieveFilters
from cadence_rag_b200.store import DenseStore, SYNTH_CORPUS_SEED, SYNTH_QUERY_SEED

rows, DIM = 1_000_000, 1024
store = DenseStore("chunks", rows, dim=DIM, device=0)
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
out = {"sync": os.environ.get("CADENCE_SYNC", "adaptive"), "cpus": len(os.sched_getaffinity(0))}
for name, f in (("filtered", filt), ("unfiltered", None)):
    for i in range(10):
        retrieve.retrieve_ids(eng, f"warm TK-{i}", f)
    t0 = time.perf_counter()
    for i in range(200):
        retrieve.retrieve_ids(eng, f"status of TK-{i % 500} and TK-{(i * 13) % 900}", f)
    out[f"{name}_1_thread"] = round(200 / (time.perf_counter() - t0))
    for n_threads in (4, 8, 16):
        per = 60
        def client(t):
            with torch.cuda.stream(torch.cuda.Stream()):
                for i in range(per):
                    retrieve.retrieve_ids(eng, f"status of TK-{(t * 97 + i) % 500} and TK-{(i * 13 + t) % 900}", f)
        ths = [threading.Thread(target=client, args=(t,)) for t in range(n_threads)]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for th in ths: th.start()
        for th in ths: th.join()
        out[f"{name}_{n_threads}_threads"] = round(n_threads * per / (time.perf_counter() - t0))
print(json.dumps(out))
