"""Shared parity assertions: CUDA top-k vs oracle top-k on identical bf16-rounded inputs.

Bar (BASELINE.md section 5 / north_star): ids identical except near-ties, scores within 1e-2
relative of the fp32 reference (we assert a far tighter 2e-5, the observed error is ~1e-6).
A "near-tie" is an oracle score gap <= TIE_REL relative: positions may swap inside such a
group, and a doc may drop out at the k-th boundary only if it near-ties the k-th score.
"""
import numpy as np

SCORE_RTOL = 2e-5      # asserted
SCORE_RTOL_SPEC = 1e-2  # what north_star allows
TIE_REL = 1e-5


def assert_topk_parity(got_scores, got_ids, all_scores, k, tie_rel=TIE_REL, rtol=SCORE_RTOL, id_offset=0,
                       scale_floor=0.0):
    """got_* [Q,k] from the CUDA path; all_scores [Q,N] fp32 oracle scores of every doc.  Tolerances are relative to
    the row's largest |score|; ``scale_floor`` (default none) bounds that scale from below for degenerate rows - a
    one-doc shard whose only score is a near-zero cancellation of large per-field terms."""
    got_scores = np.asarray(got_scores, dtype=np.float64)
    got_ids = np.asarray(got_ids, dtype=np.int64) - id_offset
    all_scores = np.asarray(all_scores, dtype=np.float64)
    Q, N = all_scores.shape
    assert got_scores.shape == (Q, k) and got_ids.shape == (Q, k)
    for q in range(Q):
        s = all_scores[q]
        scale = max(1e-30, scale_floor, np.abs(s).max())
        order = np.lexsort((np.arange(N), -s))[:k]
        ids = got_ids[q]
        assert len(set(ids.tolist())) == k, f"q{q}: duplicate ids"
        assert ids.min() >= 0 and ids.max() < N, f"q{q}: id out of range"
        # scores reported == oracle score of that id
        np.testing.assert_allclose(got_scores[q], s[ids], rtol=rtol, atol=rtol * scale, err_msg=f"q{q} scores")
        # sorted descending (up to fp noise)
        assert np.all(np.diff(got_scores[q]) <= rtol * scale), f"q{q}: not sorted"
        # set: anything missing must near-tie the k-th oracle score
        kth = s[order[-1]]
        tol = tie_rel * scale
        missing = set(order.tolist()) - set(ids.tolist())
        for m in missing:
            assert s[m] - kth <= tol, f"q{q}: doc {m} (score {s[m]}) missing, k-th is {kth}"
        extra = set(ids.tolist()) - set(order.tolist())
        for e in extra:
            assert kth - s[e] <= tol, f"q{q}: doc {e} (score {s[e]}) should not be in the top-{k}"
        # order: position p holds a doc whose oracle score near-ties the oracle's p-th score
        assert np.all(np.abs(s[ids] - s[order]) <= tol + rtol * scale), f"q{q}: rank order differs beyond near-ties"
    return True


def assert_same_topk_up_to_ties(s_a, i_a, s_b, i_b, rel=TIE_REL):
    """Two CUDA runs of the same search whose scores may differ in the last bits (fp32 atomics of the BM25 scatter
    add in a run-dependent order): rank-wise scores agree within ``rel`` of the row's scale, and an id that sits at a
    different rank (or dropped out at the k-th boundary) must near-tie the score found at that rank / the k-th."""
    s_a, s_b = np.asarray(s_a, dtype=np.float64), np.asarray(s_b, dtype=np.float64)
    i_a, i_b = np.asarray(i_a), np.asarray(i_b)
    assert s_a.shape == s_b.shape == i_a.shape == i_b.shape
    if s_a.ndim == 1:
        s_a, s_b, i_a, i_b = s_a[None], s_b[None], i_a[None], i_b[None]
    for q in range(s_a.shape[0]):
        tol = rel * max(1e-30, np.abs(s_a[q]).max())
        np.testing.assert_allclose(s_a[q], s_b[q], rtol=0, atol=tol, err_msg=f"q{q}: rank-wise scores")
        pos_b = {int(d): p for p, d in enumerate(i_b[q])}
        for p in np.nonzero(i_a[q] != i_b[q])[0]:
            d = int(i_a[q, p])
            if d in pos_b:
                assert abs(s_b[q, pos_b[d]] - s_a[q, p]) <= tol, f"q{q}: doc {d} moved between non-tied ranks"
            else:
                assert s_a[q, p] - s_a[q, -1] <= tol, f"q{q}: doc {d} missing without tying the k-th score"
