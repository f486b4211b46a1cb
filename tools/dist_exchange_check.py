"""torchrun --nproc-per-node N tools/dist_exchange_check.py : the real multi-GPU exchange paths against each other.
complete fused exchange == NCCL all_gather + merge == pipelined exchange (one call later), several epochs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multifield-adaptive-retrieval_b200"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    from mfar_b200.dist import PeerExchange, all_gather_keys, encode_keys, merge_keys
    Q, k = 37, 100
    ex = PeerExchange(q_cap=64, k_cap=128, device=dev)
    g = np.random.RandomState(100 + rank)
    prev = None
    for epoch in range(7):
        keys = torch.from_numpy(encode_keys(g.standard_normal((Q, k)).astype(np.float32),
                                            g.permutation(1000000)[: Q * k].reshape(Q, k) + 1000000 * rank
                                            ).view(np.int64)).to(dev)
        want = merge_keys(all_gather_keys(keys), k)
        if epoch % 2 == 0:
            got = ex.merge(keys, k)
            assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]), ("complete", epoch)
            prev = None
        else:
            ex.push(keys)
            got = ex.wait_merge(k, lag=1)
            torch.cuda.synchronize()
            if prev is not None:
                assert torch.equal(got[0], prev[0]) and torch.equal(got[1], prev[1]), ("pipelined", epoch)
            last = ex.wait_merge(k, lag=0)
            assert torch.equal(last[0], want[0]) and torch.equal(last[1], want[1]), ("flush", epoch)
            prev = want
        torch.cuda.synchronize()
        dist.barrier()
    # back-to-back pipelined pushes without host syncs (ranks drift apart)
    wants, gots = [], []
    for epoch in range(12):
        keys = torch.from_numpy(encode_keys(g.standard_normal((Q, k)).astype(np.float32),
                                            g.permutation(1000000)[: Q * k].reshape(Q, k) + 1000000 * rank
                                            ).view(np.int64)).to(dev)
        wants.append(merge_keys(all_gather_keys(keys), k))
        ex.push(keys)
        gots.append(ex.wait_merge(k, lag=1))
    gots.append(ex.wait_merge(k, lag=0))
    torch.cuda.synchronize()
    for e in range(12):
        assert torch.equal(gots[e + 1][0], wants[e][0]) and torch.equal(gots[e + 1][1], wants[e][1]), ("stream", e)
    if rank == 0:
        print("dist_exchange_check: ok", world, "ranks")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
