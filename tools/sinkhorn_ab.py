"""A/B timing of the persistent Sinkhorn kernel's register tiling (256 vs 512 threads per block) on the headline shape:
6 blocks of 128 x 128, lambda = 500, T = 100 and 500.  CUDA events, 5 warm-up + 50 timed launches per setting."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from otgan_b200 import _lib  # noqa: E402
from otgan_b200.utils import matching as M  # noqa: E402
from oracle import matching_oracle as mo  # noqa: E402  (input generator only)


def main():
    lib = _lib.load()
    A = torch.from_numpy(mo.synth_embeddings(256, 32768, 1, "clustered", sigma=1.0)).cuda()
    B = torch.from_numpy(mo.synth_embeddings(256, 32768, 2, "clustered", sigma=1.0)).cuda()
    a1, a2, b1, b2 = A[:128], A[128:], B[:128], B[128:]
    L = M.cost_blocks([a1, b2, a1, a1, a2, a2], [a2, b1, b1, b2, b1, b2], 500.0)
    out = {}
    for T in (100, 500):
        for rows in (4, 2):
            lib.otgan_sinkhorn_set_tile_rows(rows)
            for _ in range(5):
                M.sinkhorn(L, 500.0, T)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                M.sinkhorn(L, 500.0, T)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 50 * 1e3
            out["T%d_rows%d" % (T, rows)] = {"us": us, "sinkhorn_iters_per_sec": T / (us * 1e-6)}
    lib.otgan_sinkhorn_set_tile_rows(4)
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sinkhorn_ab.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
