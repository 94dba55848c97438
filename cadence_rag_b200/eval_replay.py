"""Replay a gold query set through the resident engine and score it the way the reference's offline
evaluation does (`eval/run_eval.py`), without Postgres.

Two halves:

* :func:`replay` sends every gold query through `retrieve_ids_batch` (the fused GPU call, 64 requests per
  call by default) and returns / writes the rows `eval/run_eval.py:78-85` reads
  (`{"query_id": ..., "retrieved_ids": ["chunk:17", "artifact_chunk:3", ...]}`), so the reference's own
  `python eval/run_eval.py --gold gold.jsonl --results results.jsonl` runs on them unchanged;
* :func:`compute_metrics` is the same arithmetic as `eval/run_eval.py:26-65` (recall@k, MRR, nDCG@k with
  binary relevance; queries without relevant ids are skipped; means over the counted queries).  The floats it
  returns are compared with the reference function's in `tests/test_host_logic.py` (golden:
  `tests/golden/reference_eval_metrics.json`).

Gold rows: `{"query_id": str, "query": str, "relevant_ids": [str], "filters": {...}?}`; `filters` takes the
`RetrieveFilters` fields (`app/schemas.py`), absent = unscoped.
"""
from __future__ import annotations

import argparse
import json
import math
from typing import Any, Dict, Iterable, List, Mapping, Optional, Sequence, Tuple

DEFAULT_KS = (5, 10, 20)            # eval/run_eval.py:72


def load_jsonl(path: str) -> List[Dict[str, Any]]:
    with open(path, "r", encoding="utf-8") as fh:
        return [json.loads(line) for line in (raw.strip() for raw in fh) if line]


def dump_jsonl(rows: Iterable[Mapping[str, Any]], path: str) -> None:
    with open(path, "w", encoding="utf-8") as fh:
        for row in rows:
            fh.write(json.dumps(row) + "\n")


def _discounted_gain(flags: Sequence[int]) -> float:
    """sum(rel / log2(rank + 1)) over ranks from 1, added in rank order (`eval/run_eval.py:18-23`)."""
    total = 0.0
    for rank, rel in enumerate(flags, start=1):
        if rel > 0:
            total += rel / math.log2(rank + 1)
    return total


def _query_scores(relevant: Sequence[str], retrieved: Sequence[str], ks: Sequence[int]) -> List[Tuple[str, float]]:
    """(metric name, value) contributions of one query; a k listed twice contributes twice, as in the reference."""
    wanted = set(relevant)
    first_hit = next((rank for rank, doc in enumerate(retrieved, start=1) if doc in wanted), 0)
    out = [("mrr", 1.0 / first_hit if first_hit else 0.0)]
    for k in ks:
        flags = [1 if doc in wanted else 0 for doc in retrieved[:k]]
        out.append((f"recall@{k}", sum(flags) / max(len(relevant), 1)))
        ideal = _discounted_gain([1] * min(len(relevant), k))
        out.append((f"ndcg@{k}", _discounted_gain(flags) / (ideal or 1.0)))
    return out


def compute_metrics(gold: Mapping[str, Sequence[str]], results: Mapping[str, Sequence[str]],
                    ks: Sequence[int] = DEFAULT_KS) -> Dict[str, float]:
    """Means of recall@k / MRR / nDCG@k over the gold queries that have at least one relevant id; a query missing
    from `results` scores 0 (`eval/run_eval.py:26-65`).  Per-key sums run in gold order, then one division."""
    sums = {name: 0.0 for name in [f"recall@{k}" for k in ks] + ["mrr"] + [f"ndcg@{k}" for k in ks]}
    counted = 0
    for query_id, relevant in gold.items():
        if not relevant:
            continue
        counted += 1
        for name, value in _query_scores(relevant, results.get(query_id, []), ks):
            sums[name] += value
    if counted == 0:
        return sums
    return {name: total / counted for name, total in sums.items()}


def _filters_of(row: Mapping[str, Any]):
    from .retrieve import RetrieveFilters
    spec = row.get("filters")
    if not spec:
        return None
    return RetrieveFilters(**spec)


def replay(engine, gold_rows: Sequence[Mapping[str, Any]], batch: int = 64,
           out_path: Optional[str] = None) -> List[Dict[str, Any]]:
    """Run every gold query through the engine's ids-only hybrid path, `batch` requests per fused call."""
    from .retrieve import retrieve_ids_batch
    results: List[Dict[str, Any]] = []
    for lo in range(0, len(gold_rows), max(1, batch)):
        part = gold_rows[lo:lo + max(1, batch)]
        responses = retrieve_ids_batch(engine, [str(r.get("query", "")) for r in part],
                                       [_filters_of(r) for r in part])
        for row, resp in zip(part, responses):
            results.append({"query_id": row["query_id"], "retrieved_ids": list(resp["retrieved_ids"])})
    if out_path:
        dump_jsonl(results, out_path)
    return results


def evaluate(engine, gold_rows: Sequence[Mapping[str, Any]], ks: Sequence[int] = DEFAULT_KS, batch: int = 64,
             out_path: Optional[str] = None) -> Dict[str, float]:
    results = replay(engine, gold_rows, batch=batch, out_path=out_path)
    gold = {row["query_id"]: list(row.get("relevant_ids", [])) for row in gold_rows}
    got = {row["query_id"]: row["retrieved_ids"] for row in results}
    return compute_metrics(gold, got, ks)


def main(argv: Optional[Sequence[str]] = None) -> None:
    """Score a results file against a gold file (the reference CLI's arguments, `eval/run_eval.py:68-73`)."""
    ap = argparse.ArgumentParser(description="Score retrieval results (recall@k, MRR, nDCG@k).")
    ap.add_argument("--gold", required=True)
    ap.add_argument("--results", required=True)
    ap.add_argument("--k", nargs="+", type=int, default=list(DEFAULT_KS))
    args = ap.parse_args(argv)
    gold = {r["query_id"]: r.get("relevant_ids", []) for r in load_jsonl(args.gold)}
    got = {r["query_id"]: r.get("retrieved_ids", r.get("retrieved", [])) for r in load_jsonl(args.results)}
    print(json.dumps(compute_metrics(gold, got, args.k), indent=2))


if __name__ == "__main__":
    main()
