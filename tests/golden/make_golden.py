"""Generates tests/golden/*.npz from the fp64 numpy oracle (oracle/matching_oracle.py).

The reference ships no golden vectors and cannot run here (TensorFlow 1.x), so these fixtures pin the ORACLE against
silent drift (and give the GPU tests committed fp64 targets); they are not reference outputs.  Re-run:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import matching_oracle as mo  # noqa: E402

CASES = {
    # name: (kind, N, D, G, lam, T, embedding kind)
    "two_batch_small": ("two_batch", 16, 64, 4, 500.0, 20, "clustered"),
    "two_batch_ragged": ("two_batch", 12, 37, 2, 100.0, 7, "iid"),
    "two_batch_T0": ("two_batch", 8, 32, 2, 500.0, 0, "iid"),
    "single_batch_small": ("single_batch", 12, 48, 3, 500.0, 15, "clustered"),
    "toy_euclid": ("toy", 16, 16, 1, 50.0, 10, "gauss"),
}


def make(name):
    kind, N, D, G, lam, T, emb = CASES[name]
    if emb == "gauss":  # notebook-2 style 2-D-ish Gaussian features (not normalised)
        rng = np.random.RandomState(7)
        A = rng.randn(N, D).astype(np.float32)
        B = (rng.randn(N, D) * 0.5 + 1.0).astype(np.float32)
    else:
        A = mo.synth_embeddings(N, D, 1, emb, sigma=0.3)
        B = mo.synth_embeddings(N, D, 2, emb, sigma=0.3)
    out = {"A": A, "B": B, "lam": lam, "T": T, "G": G}
    if kind == "two_batch":
        fa, fb = list(np.split(A, G)), list(np.split(B, G))
        res, plans, dists = mo.get_matched_features(fa, fb, lam, T, np.float64, True)
        dist = mo.calc_distance(fa, fb, res)
        ga, gb = mo.grad_features(res)
        out.update(P=np.stack(plans), C=np.stack(dists), dist=dist, grad_a=np.concatenate(ga), grad_b=np.concatenate(gb))
    elif kind == "single_batch":
        fa, fb = list(np.split(A, G)), list(np.split(B, G))
        res, plans, dists = mo.get_matched_features_single_batch(fa, fb, lam, T, np.float64, True)
        dist = mo.calc_distance(fa, fb, res)
        out.update(P=np.stack(plans), C=np.stack(dists), dist=dist)
    else:
        res, plans, dists = mo.toy_get_matched_features(A, B, lam, T, np.float64, True)
        dist = mo.toy_calc_distance(A, B, res)
        out.update(P=np.stack(plans), C=np.stack(dists), dist=dist)
    cat = (lambda x: np.concatenate(x)) if kind != "toy" else (lambda x: x)
    out.update(f_aa=cat(res[0]), f_bb=cat(res[1]), f_ab=cat(res[2]), f_ba=cat(res[3]), entropy=res[4])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return out


if __name__ == "__main__":
    for n in CASES:
        o = make(n)
        print(n, "dist=%.9g entropy=%.9g" % (o["dist"], o["entropy"]))
