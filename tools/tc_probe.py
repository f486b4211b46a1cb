"""First-contact probe of the tcgen05 kernel: one tile, F=1, w=1, k=128=N so the top-k output is a
full dump of D[128 docs x Q]; compares with the exact fp32 product and prints where it differs."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multifield-adaptive-retrieval_b200"))
from mfar_b200.modeling.retrieval import MultiFieldRetriever, PackedCorpus  # noqa: E402
from mfar_b200.modeling.weighting import LinearWeights  # noqa: E402


def probe(N, d, F, Q, impl, seed=0):
    g = torch.Generator().manual_seed(seed)
    fields = [torch.randn(N, d, generator=g).bfloat16().float() for _ in range(F)]
    q = torch.randn(Q, d, generator=g).bfloat16().float()
    pc = PackedCorpus.from_fields(fields, "cuda")
    layer = LinearWeights(F, 1).cuda()
    r = MultiFieldRetriever(pc, layer, top_k=min(N, 128), impl=impl)
    k = min(N, 128)
    s, i = r.search(q.cuda(), top_k=k)
    torch.cuda.synchronize()
    ref = sum((q @ f.t()) / F for f in fields)                  # [Q,N], uniform softmax weights
    rs, ri = torch.topk(ref, k, dim=1)
    s, i = s.cpu(), i.cpu()
    got = torch.full((Q, N), float("nan"))
    for qq in range(Q):
        valid = i[qq] >= 0
        got[qq, i[qq][valid]] = s[qq][valid]
    err = (got - ref).abs()
    finite = torch.isfinite(err)
    print(f"[{impl}] N={N} d={d} F={F} Q={Q}: covered={finite.float().mean():.3f} "
          f"max_abs_err={err[finite].max().item() if finite.any() else float('nan'):.3e} "
          f"ids_match={(i == ri).float().mean():.3f} ref_scale={ref.abs().max():.2f}")
    if finite.any() and err[finite].max() > 1e-2:
        bad = (err > 1e-2) & finite
        print("   bad per query:", bad.sum(1).tolist()[:16])
        print("   bad per doc (first 32 docs):", bad.sum(0).tolist()[:32])
        print("   got[0,:8]", got[0, :8].tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())
    return bool(finite.any() and err[finite].max() < 1e-2 and (i == ri).float().mean() > 0.99)


if __name__ == "__main__":
    ok = True
    for impl in sys.argv[1:] or ["simt", "tcgen05"]:
        shapes = [(128, 64, 1, 1), (128, 64, 1, 16), (128, 128, 1, 16), (128, 768, 1, 16),
                  (128, 768, 2, 16), (256, 768, 3, 40), (1000, 768, 4, 64), (5000, 768, 8, 3)]
        if impl == "tcgen05_qs":
            shapes += [(128, 64, 1, 128), (1000, 768, 2, 100), (128, 64, 1, 130), (128, 768, 1, 256),
                       (3000, 768, 3, 256), (40000, 768, 8, 512), (20000, 768, 2, 700)]
        for (N, d, F, Q) in shapes:
            try:
                ok &= probe(N, d, F, Q, impl)
            except Exception as e:  # noqa: BLE001
                print(f"[{impl}] N={N} d={d} F={F} Q={Q}: EXCEPTION {e}")
                ok = False
                break
    print("PROBE", "OK" if ok else "FAILED")
