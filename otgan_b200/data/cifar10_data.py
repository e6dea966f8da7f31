"""Minimal stand-in for the reference's data/cifar10_data.py:40 `load(data_dir, subset)`: reads the CIFAR-10 python
pickles if they are present (no download -- there is no network), else raises FileNotFoundError.  `synthetic(n)` gives
CIFAR-10-shaped uniform noise for benchmarks (BASELINE.json configs are synthetic)."""
import os
import pickle

import numpy as np


def _unpickle(path):
    with open(path, "rb") as fo:
        d = pickle.load(fo, encoding="latin1")
    return {"x": d["data"].reshape((10000, 3, 32, 32)), "y": np.array(d["labels"]).astype(np.uint8)}


def load(data_dir, subset="train"):
    base = os.path.join(data_dir, "cifar-10-batches-py")
    if subset == "train":
        parts = [_unpickle(os.path.join(base, "data_batch_%d" % i)) for i in range(1, 6)]
        return np.concatenate([p["x"] for p in parts], 0), np.concatenate([p["y"] for p in parts], 0)
    if subset == "test":
        p = _unpickle(os.path.join(base, "test_batch"))
        return p["x"], p["y"]
    raise NotImplementedError("subset should be either train or test")


def synthetic(n, seed=1, size=32):
    """[n, size, size, 3] float32 in [-1, 1] (already NHWC and scaled like train.py:158)."""
    rng = np.random.RandomState(seed)
    return (rng.rand(n, size, size, 3).astype(np.float32) * 2.0 - 1.0)
