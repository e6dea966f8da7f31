"""Evaluation hook of the reference (utils/inception.py:24-54, called from train.py:247-270 every 10 epochs): the Inception score
of a list of generated images.

    get_inception_score(images, splits=10, predict_fn=None) -> (mean, std)

Same signature, input checks and arithmetic as the reference: `images` is a list of HWC numpy arrays in [0, 255]; batches of 100 go
through the classifier; the predictions are cut into `splits` parts, score_k = exp(mean_i KL(p(y|x_i) || p(y))) per part
(:44-52).  The classifier itself -- the 2015 Inception graph the reference downloads at import time (:18, :57-94) -- cannot exist
here (no network, no TensorFlow), so it is an ARGUMENT: `predict_fn(batch [n, H, W, 3] float32 in [0, 255]) -> [n, classes]`
softmax probabilities (e.g. a torchvision Inception-v3 wrapped by the caller).  Without one the call raises: there is no silent
substitute for the published metric.  Not on the benchmarked path."""
import math
import sys

import numpy as np


def score_from_predictions(preds, splits=10):
    """utils/inception.py:43-52 on an [n, classes] array of softmax outputs."""
    preds = np.asarray(preds, dtype=np.float64)
    scores = []
    for i in range(splits):
        part = preds[(i * preds.shape[0] // splits):((i + 1) * preds.shape[0] // splits), :]
        kl = part * (np.log(part) - np.log(np.expand_dims(np.mean(part, 0), 0)))
        kl = np.mean(np.sum(kl, 1))
        scores.append(np.exp(kl))
    return float(np.mean(scores)), float(np.std(scores))


def get_inception_score(images, splits=10, predict_fn=None, batch_size=100, progress=False):
    assert type(images) == list                                   # utils/inception.py:25-29
    assert type(images[0]) == np.ndarray
    assert len(images[0].shape) == 3
    assert np.max(images[0]) > 10
    assert np.min(images[0]) >= 0.0
    if predict_fn is None:
        raise RuntimeError("get_inception_score needs predict_fn: the reference's Inception-2015 graph "
                           "(download.tensorflow.org/models/image/imagenet/inception-2015-12-05.tgz) is not available offline")
    preds = []
    n_batches = int(math.ceil(float(len(images)) / float(batch_size)))
    for i in range(n_batches):
        if progress:
            sys.stdout.write(".")
            sys.stdout.flush()
        inp = np.stack([img.astype(np.float32) for img in images[(i * batch_size):min((i + 1) * batch_size, len(images))]], 0)
        preds.append(np.asarray(predict_fn(inp)))
    return score_from_predictions(np.concatenate(preds, 0), splits)
