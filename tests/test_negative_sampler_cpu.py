"""IndexNegativeSampler mirror vs what the reference's own class returned for the same fake index, positives and
``random`` seed (tests/golden/negative_sampler.json, oracle/make_golden_sampler.py).  No GPU needed."""
import json
import os
import random

from make_golden_sampler import FakeIndex, hit_table
from mfar_b200.data.negative_sampler import Document, IndexNegativeSampler, Query

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "negative_sampler.json")))


def _setup(case):
    table = hit_table(GOLD["table_seed"], GOLD["n_queries"], GOLD["n_docs"])
    index = FakeIndex(table)
    sampler = IndexNegativeSampler(index, {f"d{j}": f"text {j}" for j in range(0, 80, 2)}, n_retrieve=case["n_retrieve"],
                                   n_bottom=case["n_bottom"], n_sample=case["n_sample"])
    pos = {k: set(v) for k, v in case["pos"].items()}
    queries = [Query(str(i), q) for i, q in enumerate(table)]
    return index, sampler, pos, queries


def test_sample_matches_reference_per_query_loop():
    for case in GOLD["cases"]:
        index, sampler, pos, queries = _setup(case)
        assert sampler.n_sample == case["n_sample"]
        random.seed(case["seed"])
        got = [[[d._id, d.text] for d in sampler.sample(q, pos)] for q in queries]
        assert got == case["picked"]
        assert [list(c) for c in index.calls] == case["calls"]          # same retrievals, incl. the deeper retry


def test_sample_batch_is_one_batched_retrieval_with_the_same_result():
    for case in GOLD["cases"]:
        index, sampler, pos, queries = _setup(case)
        batched = []
        index.retrieve_batch = lambda qs, top_k, _i=index: (batched.append(len(qs)) or [_i.retrieve(q, top_k) for q in qs])
        random.seed(case["seed"])
        got = sampler.sample_batch(queries, pos)
        assert [[[d._id, d.text] for d in docs] for docs in got] == case["picked"]
        assert batched == [len(queries)]                                 # ONE retrieve_batch for the whole batch
        assert all(isinstance(d, Document) for docs in got for d in docs)


def test_corpus_container_matches_reference_accessors():
    from mfar_b200.data.typedef import Corpus
    c = Corpus.from_docs_dict({"a": "alpha text", "b": "beta"}, dataset_name="mag")
    assert list(c.keys()) == ["a", "b"] and len(c) == 2 and c.dataset_name == "mag"
    assert c.get_text_by_id(1) == "beta" and c.get_text_by_key("a") == "alpha text"
    assert c.get_doc_by_key("b")._id == "b" and list(c.pairs()) == [("a", "alpha text"), ("b", "beta")]
    try:
        c.get_doc_by_key("zzz")
        raise AssertionError("expected KeyError")
    except KeyError as e:
        assert "not found in corpus" in str(e)
