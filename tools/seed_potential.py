#!/usr/bin/env python
"""Experiment (needs a library built with -DMFAR_DEBUG_SEED, MFAR_LIB=...): how much of a shard-sized scoring kernel is
candidate-list work that a tighter admission seed would remove?  Runs the same search with (a) the prefix seed, (b) the
prefix seed replaced by the batch's TRUE k-th keys (oracle seed: the fewest admissions any seed can give), (c) the k-th
keys of an 8x larger prefix (what a seed merged over 8 ranks' prefixes would be)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multifield-adaptive-retrieval_b200"))
import torch  # noqa: E402


def main():
    docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
    Q = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    from mfar_b200 import _native as nv
    from mfar_b200 import synth
    from mfar_b200.modeling.retrieval import MultiFieldRetriever, PackedCorpus
    from mfar_b200.modeling.weighting import LinearWeights
    lib = nv.lib()
    lib.mfar_debug_set_seed.argtypes = [ctypes.c_void_p]
    lib.mfar_debug_set_seed.restype = None
    dev = torch.device("cuda", 0)
    pc = PackedCorpus(docs, 8, 768, dev)
    synth.fill_packed_corpus(pc, seed=1234)
    mu = synth.corpus_mean(768, 1234, dev)
    layer = LinearWeights(768, 8, query_cond=True)
    with torch.no_grad():
        layer.weight.copy_(synth.make_mixture(768, 8, 1235))
    r = MultiFieldRetriever(pc, layer.to(dev), top_k=100)
    q = synth.make_queries(Q, 768, mu, 1334, dev)

    def timed(tag, seed):
        lib.mfar_debug_set_seed(None if seed is None else seed.data_ptr())
        for _ in range(3):
            r.search(q, q.float())
        nv.check(lib.mfar_profile_enable(1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            out = r.search(q, q.float(), return_keys=True)
        e1.record(); e1.synchronize()
        buf = (ctypes.c_float * 256)()
        n = lib.mfar_profile_collect(ctypes.addressof(buf), 256)
        lib.mfar_profile_enable(0)
        km = sum(buf[i] for i in range(n)) / max(n, 1)
        print(json.dumps({"seed": tag, "docs": docs, "batch": Q, "step_ms": e0.elapsed_time(e1) / 20, "kernel_ms": km}))
        return out

    s0, i0, k0 = timed("prefix (shipped)", None)
    true_kth = (k0[:, -1].contiguous() - 1).contiguous()                 # exclusive, as seed_from_keys makes it
    s1, i1, _ = timed("true k-th keys (lower bound on admissions)", true_kth)
    assert torch.equal(i0, i1) and torch.equal(s0, s1)
    # the seed a merged prefix of 8 ranks would give: k-th key of the first 8 * 148 tiles
    n8 = min(docs, 8 * 148 * 128)
    r8 = MultiFieldRetriever(pc.window(0, n8), layer.to(dev), top_k=100)
    lib.mfar_debug_set_seed(None)
    _, _, k8 = r8.search(q, q.float(), return_keys=True)
    s2, i2, _ = timed("k-th keys of an 8x larger prefix", (k8[:, -1].contiguous() - 1).contiguous())
    assert torch.equal(i0, i2) and torch.equal(s0, s2)
    lib.mfar_debug_set_seed(None)


if __name__ == "__main__":
    main()
