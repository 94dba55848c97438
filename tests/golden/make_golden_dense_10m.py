"""Generates tests/golden/dense_10m.json: the C oracle's top-51 (both variants) of headline queries 0..3 over the
full 10 M x 1024 synthetic corpus (BASELINE configs[2] / north-star size), streamed in 250 000-row chunks.
~3 minutes on 8 cores.  The GPU tier compares the engine with these lists without repeating the 41 GB CPU pass.

    python tests/golden/make_golden_dense_10m.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu_oracle as orc  # noqa: E402

N, K, CHUNK = 10_000_000, 50, 250_000
CORPUS_SEED, QUERY_SEED = 20260209, 20260210


def main():
    nq = 4
    qs = orc.synth_rows(QUERY_SEED, 0, nq)
    variants = {"f64": orc.VARIANT_F64, "pgv32": orc.VARIANT_PGV32}
    best = [{name: [] for name in variants} for _ in range(nq)]
    for r0 in range(0, N, CHUNK):
        x = orc.synth_rows(CORPUS_SEED, r0, CHUNK)
        rid = np.arange(r0 + 1, r0 + CHUNK + 1, dtype=np.int64)
        for qi in range(nq):
            for name, v in variants.items():
                best[qi][name].append(orc.exact_scan(qs[qi], x, K + 14, ids=rid, variant=v))
        print(f"rows {r0 + CHUNK}", flush=True)
    out = {"rows": N, "dim": 1024, "k": K, "corpus_seed": CORPUS_SEED, "query_seed": QUERY_SEED,
           "note": "top-51 of the C oracle (oracle/pgvector_restated.c) per query row and variant; scores as float.hex()",
           "queries": []}
    for qi in range(nq):
        rec = {"query_row": qi}
        for name, acc in best[qi].items():
            ids = np.concatenate([a for a, _ in acc]); sc = np.concatenate([b for _, b in acc])
            order = np.lexsort((ids, -sc))[:K + 1]
            rec[name] = {"ids": ids[order].tolist(), "scores_hex": [float(v).hex() for v in sc[order]]}
        sc = np.array([float.fromhex(h) for h in rec["f64"]["scores_hex"]])
        gaps = np.abs(np.diff(sc)) / np.abs(sc[:-1])
        rec["min_relative_gap_top51"] = float(gaps.min())
        rec["min_gap_position"] = int(gaps.argmin())
        rec["variants_agree"] = rec["f64"]["ids"] == rec["pgv32"]["ids"]
        out["queries"].append(rec)
        print(qi, "min relative gap", rec["min_relative_gap_top51"], "at", rec["min_gap_position"], "variants agree:", rec["variants_agree"])
    with open(os.path.join(ROOT, "tests", "golden", "dense_10m.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
