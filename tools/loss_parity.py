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


def trajectory(backend, steps, n, t_iters, lam, model="dcgan", perturb=0.0):
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
            if perturb:                                  # control run: rounding-level noise on the inputs of every step
                x = x * (1.0 + perturb * torch.randn(x.shape, device="cuda", generator=None))
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
    """A: tcgen05 (TF32 operands); B: strict fp32 library rung; C: B again with 1e-6 relative noise on the input images -- the
    CONTROL that shows how fast this (chaotic) GAN trajectory separates from itself under fp32-rounding-level perturbations."""
    a = trajectory("tcgen05", steps, n, t_iters, lam, model)
    b = trajectory("cudnn", steps, n, t_iters, lam, model)
    c = trajectory("cudnn", steps, n, t_iters, lam, model, perturb=1e-6)

    def gaps(p, q):
        dd = [abs(x[1] - y[1]) for x, y in zip(p, q)]
        de = [abs(x[2] - y[2]) for x, y in zip(p, q)]
        win = lambda v, lo, hi: sum(v[lo:hi]) / max(1, len(v[lo:hi]))
        return {"distance_gap_mean_steps_0_20": win(dd, 0, 20), "distance_gap_mean_steps_20_60": win(dd, 20, 60),
                "distance_gap_mean_rest": win(dd, 60, steps), "distance_gap_max_steps_0_20": max(dd[:20]),
                "entropy_gap_mean_steps_0_20": win(de, 0, 20), "entropy_gap_mean_steps_20_60": win(de, 20, 60),
                "entropy_gap_mean_rest": win(de, 60, steps), "entropy_gap_max_steps_0_20": max(de[:20])}

    mean = lambda t, i, lo: sum(x[i] for x in t[lo:]) / max(1, len(t[lo:]))
    return {"model": model, "steps": steps, "N": n, "T": t_iters, "lambda": lam,
            "distance_scale_fp32": max(abs(y[1]) for y in b),
            "tf32_vs_fp32": gaps(a, b), "fp32_vs_fp32_perturbed_1e-6": gaps(b, c),
            "late_means": {"tf32": [mean(a, 1, steps // 2), mean(a, 2, steps // 2)], "fp32": [mean(b, 1, steps // 2), mean(b, 2, steps // 2)],
                           "fp32_perturbed": [mean(c, 1, steps // 2), mean(c, 2, steps // 2)]},
            "every_10th": [(s, a[s][1], b[s][1], c[s][1], a[s][2], b[s][2], c[s][2]) for s in range(0, steps, 10)]}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--T", type=int, default=100)
    ap.add_argument("--lam", type=float, default=500.0)
    ap.add_argument("--model", default="dcgan")
    a = ap.parse_args()
    print(json.dumps(compare(a.steps, a.n, a.T, a.lam, a.model)))
