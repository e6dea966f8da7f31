"""One launch of each matching kernel at the headline size (N = 256, D = 32768, T = 100): the target of the ncu captures.

    ncu --set full --import-source on --clock-control none -o gpurun_out/prof python tools/matching_once.py [reps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otgan_b200 import _lib                      # noqa: E402
from otgan_b200.utils import matching as M       # noqa: E402
from oracle import matching_oracle as mo          # noqa: E402  (input generator only)


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    dev = torch.device("cuda", 0)
    N, D, lam, T = 256, 32768, 500.0, 100
    A = torch.from_numpy(mo.synth_embeddings(N, D, 100, "clustered", sigma=1.0)).to(dev)
    B = torch.from_numpy(mo.synth_embeddings(N, D, 101, "clustered", sigma=1.0)).to(dev)
    fa, fb = list(torch.chunk(A, 2, 0)), list(torch.chunk(B, 2, 0))
    for _ in range(reps):
        M.matching_step(fa, fb, lam, T)
    torch.cuda.synchronize()
    print("launches:", _lib.launch_count())


if __name__ == "__main__":
    main()
