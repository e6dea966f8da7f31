"""Worker of tests/test_train_host.py::test_two_rank_gather_and_grad_sum_gloo (one process per rank, gloo, CPU)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otgan_b200 import train as T  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bs, D = 3, 5
    f_gen = torch.full((bs, D), 10.0 * rank) + torch.arange(bs).float()[:, None]
    f_dat = -f_gen
    A, B = T.gather_features(f_gen, f_dat, world)
    lo, hi = T.local_rows(rank, bs)
    ok = A.shape == (world * bs, D) and torch.equal(A[lo:hi], f_gen) and torch.equal(B[lo:hi], f_dat)
    ok = ok and torch.equal(A[:bs, 0], torch.arange(bs).float()) and torch.equal(A[bs:2 * bs, 0], 10.0 + torch.arange(bs).float())
    g = torch.full((4,), float(rank + 1))                 # summed (not averaged) tower gradients, train.py:134-139
    dist.all_reduce(g, op=dist.ReduceOp.SUM)
    ok = ok and torch.equal(g, torch.full((4,), float(sum(range(1, world + 1)))))
    dist.destroy_process_group()
    print("RANK %d %s" % (rank, "OK" if ok else "FAIL"))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
