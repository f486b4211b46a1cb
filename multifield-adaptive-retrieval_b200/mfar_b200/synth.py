"""Synthetic STaRK-shaped inputs, seeded, generated where they are used.

corpus  x = z + 0.5*mu, z ~ N(0,1) iid, mu ~ N(0,1)^d shared by the corpus (positive, clustered
        scores like real Contriever outputs), rounded to bf16 ONCE - that bf16 tensor is the
        input of both the CUDA path and the CPU oracle.
queries same distribution;  W ~ 0.05*N(0,1) [d,F];  sparse: 0 w.p. 0.95 else Gamma(2,2), fp16.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

SHAPES = {
    # name: (n_docs, n_dense, n_sparse)   d = 768 everywhere (BASELINE.json configs)
    "prime_2k": (2000, 22, 0),
    "prime_full": (129375, 22, 22),
    "mag_full": (700244, 5, 0),
    "amazon_full": (957192, 8, 8),
    "scale_10m_all": (10_000_000, 8, 0),
    "scale_10m_single": (10_000_000, 1, 0),
    # sparse-only scorers (all_sparse field sets): no dense contraction, the pass is BM25 + streaming top-k
    "amazon_sparse": (957192, 0, 8),
    "prime_sparse": (129375, 0, 22),
}


def _gen(device, seed: int) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def corpus_mean(dim: int, seed: int, device="cpu") -> torch.Tensor:
    return torch.randn(dim, generator=_gen("cpu", seed), dtype=torch.float32).to(device)


def make_field_rows(n_rows: int, dim: int, mu: torch.Tensor, gen: torch.Generator, device) -> torch.Tensor:
    z = torch.randn((n_rows, dim), generator=gen, dtype=torch.float32, device=device)
    return (z + 0.5 * mu).to(torch.bfloat16)


def make_queries(Q: int, dim: int, mu: torch.Tensor, seed: int, device="cpu") -> torch.Tensor:
    g = _gen(device, seed)
    z = torch.randn((Q, dim), generator=g, dtype=torch.float32, device=device)
    return (z + 0.5 * mu.to(device)).to(torch.bfloat16)


def make_mixture(dim: int, F: int, seed: int, query_cond: bool = True, device="cpu") -> torch.Tensor:
    g = _gen("cpu", seed)
    if query_cond:
        return (0.05 * torch.randn((dim, F), generator=g)).to(device)
    return torch.randn((F, 1), generator=g).to(device)


def make_sparse(Q: int, Fs: int, n_docs: int, seed: int, device="cpu", dtype=torch.float16,
                pitch: int = 1) -> Optional[torch.Tensor]:
    """[Q, Fs, ld] with ld = n_docs rounded up to ``pitch`` (columns beyond n_docs are zero)."""
    if Fs == 0:
        return None
    if pitch > 1 and n_docs % pitch:
        ld = (n_docs + pitch - 1) // pitch * pitch
        out = torch.zeros((Q, Fs, ld), dtype=dtype, device=device)
        out[:, :, :n_docs] = make_sparse(Q, Fs, n_docs, seed, device, dtype)
        return out
    g = _gen(device, seed)
    u = torch.rand((Q, Fs, n_docs), generator=g, device=device)
    # Gamma(2, scale 2) = sum of two Exp(scale 2)
    e1 = -2.0 * torch.log(torch.rand((Q, Fs, n_docs), generator=g, device=device).clamp_min(1e-12))
    e2 = -2.0 * torch.log(torch.rand((Q, Fs, n_docs), generator=g, device=device).clamp_min(1e-12))
    return torch.where(u < 0.95, torch.zeros_like(e1), e1 + e2).to(dtype)


# ---- synthetic BM25 fields (device-resident sparse scorer): token ids ~ Zipf(1) over a 30k vocabulary, field length
# ~ Poisson(12) tokens; query tokens are Zipf-distributed over ranks > 20 (stop-word-like head terms never appear in
# queries), 8 per (query, field): the union of their postings covers ~5 % of the docs - the density make_sparse uses.
BM25_VOCAB = 30000
BM25_DOC_LEN = 12
BM25_QUERY_TOKENS = 8
BM25_QUERY_MIN_RANK = 20


def make_bm25_field(n_docs: int, seed: int, device, n_vocab: int = BM25_VOCAB, mean_len: int = BM25_DOC_LEN,
                    doc_range: Optional[Tuple[int, int]] = None):
    """One sparse field of the WHOLE corpus indexed on the device (idf / average length are corpus statistics), then
    cut to ``doc_range``.  Deterministic in (n_docs, seed) so every shard count sees the same global field."""
    from .data.bm25 import DeviceBM25
    g = _gen(device, seed)
    p = 1.0 / torch.arange(1, n_vocab + 1, dtype=torch.float32, device=device)
    lens = torch.poisson(torch.full((n_docs,), float(mean_len), device=device), generator=g).clamp_(min=1).long()
    total = int(lens.sum().item())
    toks = torch.empty(total, dtype=torch.int64, device=device)
    step = 1 << 24                                           # torch.multinomial draws at most 2^24 samples per call
    for lb in range(0, total, step):
        n = min(step, total - lb)
        toks[lb:lb + n] = torch.multinomial(p, n, replacement=True, generator=g)
    full = DeviceBM25(device=device).index_flat(toks, lens, n_vocab)
    return full.shard(*doc_range) if doc_range is not None else full


def make_bm25_query_entries(Q: int, Fs: int, seed: int, n_vocab: int = BM25_VOCAB,
                            n_tokens: int = BM25_QUERY_TOKENS) -> torch.Tensor:
    """HOST int32 [Q*Fs*n_tokens, 3] entries (query row, sparse field, token id), query-major like
    ``data.bm25.token_entries``."""
    g = _gen("cpu", seed)
    p = 1.0 / torch.arange(1, n_vocab + 1, dtype=torch.float32)
    p[:BM25_QUERY_MIN_RANK] = 0
    tok = torch.multinomial(p, Q * Fs * n_tokens, replacement=True, generator=g).view(Fs, Q, n_tokens)
    ent = torch.empty((Fs, Q, n_tokens, 3), dtype=torch.int32)
    ent[..., 0] = torch.arange(Q, dtype=torch.int32).view(1, Q, 1)
    ent[..., 1] = torch.arange(Fs, dtype=torch.int32).view(Fs, 1, 1)
    ent[..., 2] = tok.int()
    return ent.permute(1, 0, 2, 3).reshape(-1, 3).contiguous()


def fill_packed_corpus(pc, seed: int, chunk_docs: int = 65536, doc_offset: int = 0) -> None:
    """Generate the corpus ON DEVICE straight into a PackedCorpus (config 5 is 122.9 GB: it cannot come
    from host RAM).  Doc n's vectors depend only on (seed, global chunk index), so shards of a
    doc-sharded run see the same global corpus when ``doc_offset`` is a multiple of ``chunk_docs``."""
    assert doc_offset % chunk_docs == 0
    mu = corpus_mean(pc.dim, seed, pc.device)
    for lb in range(0, pc.n_docs, chunk_docs):
        ub = min(pc.n_docs, lb + chunk_docs)
        g = _gen(pc.device, seed * 1000003 + (doc_offset + lb) // chunk_docs + 1)
        for f in range(pc.n_fields):
            pc.load_rows(f, lb, make_field_rows(ub - lb, pc.dim, mu, g, pc.device))
