"""ncu target: fused hybrid calls (64 unscoped requests, tech tokens, dense lane on the bf16 tensor-core lane)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from cadence_rag_b200 import _ffi, embeddings, retrieve
from cadence_rag_b200.config import settings
from cadence_rag_b200.lexical import TechTokenIndex
from cadence_rag_b200.retrieve import DenseEngine
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
emb = embeddings.SyntheticEmbedder(seed=SYNTH_QUERY_SEED, dim=DIM)
settings.embeddings_dim = DIM
dev_index = eng.device_tech_indexes["chunks"]
B = 64
texts = [f"status of TK-{i % 500} and TK-{(i * 13) % 900}" for i in range(B)]
qv = np.stack([np.asarray(emb([t]).vectors[0], dtype=np.float32) for t in texts])
tk, nt = dev_index.encode_tokens([retrieve.extract_tech_tokens(t) for t in texts])
ann = dict(call_slots=None, date_from=None, date_to=None, tag_mask=None, dense_lane=_ffi.CDR_DENSE_LANE_BATCH_BF16)
for label, kw in (("ann+tech", dict(tech_index=dev_index, token_ids=tk, n_tokens=nt)), ("ann only", {})):
    for _ in range(3):
        store.hybrid_retrieve(qv, 50, filter_spec=ann, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        store.hybrid_retrieve(qv, 50, filter_spec=ann, **kw)
    print(label, f"{(time.perf_counter() - t0) / 10 * 1e3:.3f} ms per 64-request call")
