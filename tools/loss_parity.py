"""Loss-parity study: the same OT-GAN training trajectory (fixed seeds, images and latents) on the tcgen05 convolution kernels
(TF32 operands) and on the strict-fp32 library rung (cuDNN, allow_tf32 = False) -- the precision class of the reference's
TensorFlow-1.x fp32 convolutions.  Prints one JSON object with the per-step distance / entropy of both runs and their gaps.

    python tools/loss_parity.py [--steps 200] [--n 64] [--T 100]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otgan_b200 import train as T  # noqa: E402
from otgan_b200.utils import nn  # noqa: E402


def trajectory(backend, steps, n, t_iters, lam, model="dcgan"):
    prev = (nn.CONV_BACKEND, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    nn.CONV_BACKEND = backend
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        args = T.build_parser().parse_args(["--synthetic", "--nr_gpu", "2", "--batch_size", str(n // 2), "--nr_sinkhorn_iter", str(t_iters),
                                            "--sinkhorn_lambda", str(lam), "--model", model, "--seed", "3"])
        tr = T.Trainer(args, torch.device("cuda", 0))
        g = torch.Generator().manual_seed(99)
        out = []
        for s in range(steps):
            x = (torch.rand((n, 32, 32, 3), generator=g) * 2 - 1).cuda()
            if model == "dcgan":
                u = (torch.rand((n, 100), generator=g) * 2 - 1).cuda()
            else:
                u = [(torch.rand((n, 100), generator=g) * 2 - 1).cuda()] + \
                    [(torch.rand((n, sz, sz, 16), generator=g) * 2 - 1).cuda() for sz in (8, 16, 32)]
            kind, stats = tr.step(x, u=u)
            d, e = stats.tolist()
            out.append((kind, d, e))
        return out
    finally:
        nn.CONV_BACKEND, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def compare(steps=200, n=64, t_iters=100, lam=500.0, model="dcgan"):
    a = trajectory("tcgen05", steps, n, t_iters, lam, model)
    b = trajectory("cudnn", steps, n, t_iters, lam, model)
    dd = [abs(x[1] - y[1]) for x, y in zip(a, b)]
    de = [abs(x[2] - y[2]) for x, y in zip(a, b)]
    scale_d = max(abs(y[1]) for y in b)
    return {"model": model, "steps": steps, "N": n, "T": t_iters, "lambda": lam,
            "max_abs_gap_distance": max(dd), "mean_abs_gap_distance": sum(dd) / len(dd), "max_abs_distance_fp32": scale_d,
            "max_abs_gap_entropy": max(de), "mean_abs_gap_entropy": sum(de) / len(de),
            "final": {"tcgen05": a[-1][1:], "fp32": b[-1][1:]},
            "first_steps_gap_distance": dd[:12], "every_20th": [(s, a[s][1], b[s][1], a[s][2], b[s][2]) for s in range(0, steps, 20)]}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--T", type=int, default=100)
    ap.add_argument("--lam", type=float, default=500.0)
    ap.add_argument("--model", default="dcgan")
    a = ap.parse_args()
    print(json.dumps(compare(a.steps, a.n, a.T, a.lam, a.model)))
