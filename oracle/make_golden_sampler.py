"""Generate tests/golden/negative_sampler.json by EXECUTING the reference's own ``IndexNegativeSampler``
(mfar/data/negative_sampler.py, unmodified) over a fake ``Index`` that serves fixed hit lists.

Run in the build container only:   python oracle/make_golden_sampler.py
"""
from __future__ import annotations

import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "negative_sampler.json")


def hit_table(seed: int, n_queries: int, n_docs: int):
    """query text -> full ranking [(doc key, score)] (descending, with score ties), deterministic."""
    rng = random.Random(seed)
    table = {}
    for i in range(n_queries):
        docs = [f"d{j}" for j in range(n_docs)]
        rng.shuffle(docs)
        scores = sorted((round(rng.uniform(0, 10), 1) for _ in docs), reverse=True)      # 1 decimal: ties
        table[f"query {i}"] = list(zip(docs, scores))
    return table


class FakeIndex:
    def __init__(self, table):
        self.table = table
        self.calls = []

    def retrieve(self, query, top_k):
        self.calls.append((query, top_k))
        return self.table[query][:top_k]

    def retrieve_batch(self, queries, top_k):
        return [self.retrieve(q, top_k) for q in queries]


def main():
    ref_import.load()
    from mfar.data.negative_sampler import IndexNegativeSampler
    from mfar.data.typedef import Query
    table = hit_table(7, n_queries=12, n_docs=80)
    cases = []
    for n_retrieve, n_bottom, n_sample, seed in [(50, 5, 1, 0), (20, 8, 3, 1), (10, 5, 2, 2)]:
        pos = {}
        for i, (q, ranking) in enumerate(table.items()):
            top = [d for d, _ in ranking]
            if i % 4 == 0:
                pos[str(i)] = set(top[:n_retrieve])                   # every retrieved doc is a positive -> retry path
            elif i % 4 == 1:
                pos[str(i)] = set(top[n_retrieve - 3:n_retrieve])     # positives inside the bottom slice
            else:
                pos[str(i)] = set(top[1:4])
        index = FakeIndex(table)
        sampler = IndexNegativeSampler(index, {f"d{j}": f"text {j}" for j in range(0, 80, 2)}, n_retrieve=n_retrieve,
                                       n_bottom=n_bottom, n_sample=n_sample)
        random.seed(seed)
        queries = [Query(str(i), q) for i, q in enumerate(table)]
        picked = [[(d._id, d.text) for d in docs] for docs in sampler.sample_batch(queries, pos)]
        cases.append(dict(n_retrieve=n_retrieve, n_bottom=n_bottom, n_sample=n_sample, seed=seed,
                          pos={k: sorted(v) for k, v in pos.items()}, picked=picked, calls=index.calls))
    with open(OUT, "w") as f:
        json.dump(dict(table_seed=7, n_queries=12, n_docs=80, cases=cases), f)
    print("wrote", OUT, [len(c["calls"]) for c in cases])


if __name__ == "__main__":
    main()
