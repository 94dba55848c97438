"""Golden vectors for `cadence_rag_b200.eval_replay.compute_metrics`: outputs of the reference's own
`eval/run_eval.py:compute_metrics` (stdlib only, imported from /root/reference in the build container) on seeded
gold / result sets.  Floats are stored as hex so the comparison is bit-exact.

    python tests/golden/make_golden_eval.py        # rewrites tests/golden/reference_eval_metrics.json
"""
import importlib.util
import json
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))


def _reference():
    spec = importlib.util.spec_from_file_location("ref_run_eval", "/root/reference/eval/run_eval.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _doc(rng):
    return f"{rng.choice(['chunk', 'artifact_chunk'])}:{rng.randint(1, 400)}"


def cases():
    rng = random.Random(20260209)
    out = []
    for case in range(24):
        gold, results = {}, {}
        for qi in range(rng.randint(0, 12)):
            qid = f"q{case}_{qi}"
            relevant = [] if rng.random() < 0.15 else sorted({_doc(rng) for _ in range(rng.randint(1, 6))})
            gold[qid] = relevant
            if rng.random() < 0.1:
                continue                                     # query missing from the results file
            retrieved = [_doc(rng) for _ in range(rng.randint(0, 30))]
            for doc in relevant:                             # plant some hits at random ranks
                if rng.random() < 0.6:
                    retrieved.insert(rng.randint(0, len(retrieved)), doc)
            if rng.random() < 0.2 and retrieved:
                retrieved.append(retrieved[0])               # a duplicate id counts twice in the reference
            results[qid] = retrieved
        ks = rng.choice([[5, 10, 20], [1, 3], [50], [2, 2], [10, 5]])
        out.append({"gold": gold, "results": results, "ks": ks})
    return out


def main():
    ref = _reference()
    rows = []
    for c in cases():
        m = ref.compute_metrics(c["gold"], c["results"], c["ks"])
        rows.append({**c, "metric_order": list(m.keys()), "metrics_hex": {k: float(v).hex() for k, v in m.items()}})
    with open(os.path.join(HERE, "reference_eval_metrics.json"), "w") as f:
        json.dump({"source": "eval/run_eval.py:26-65 via tests/golden/make_golden_eval.py", "cases": rows}, f, indent=0)
    print(len(rows), "cases")


if __name__ == "__main__":
    main()
