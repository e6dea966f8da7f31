"""torch-CPU restatement of the reference's matching path, ONE torch op per TensorFlow op -- TEST / BASELINE INFRASTRUCTURE.

This is the "reference-restated" CPU baseline BASELINE.md section 3 prescribes: the same op granularity as the reference's
TensorFlow graph (tf.matmul -> torch.matmul, tf.reduce_logsumexp -> torch.logsumexp, tf.nn.softmax -> torch.softmax,
softmax_cross_entropy_with_logits -> -(p * log_softmax).sum), executed by torch's multi-threaded CPU kernels (MKL / oneDNN)
in fp32.  Only tests/ and bench.py's CPU legs import it.  Pinned against tests/golden/ref_*.npz (the reference's own code
executed over the numpy shim) in tests/test_reference_golden.py.

Reference lines followed (paths relative to /root/reference):
  get_matched_features   utils/matching.py:11-85     calc_distance   utils/matching.py:139-153
"""
import time

import torch


def get_matched_features(features_a, features_b, sinkhorn_lambda, nr_sinkhorn_iter, phases=None):
    """utils/matching.py:11-85 on lists of per-tower torch CPU tensors.  `phases` (optional list) receives the wall
    seconds of the cost / Sinkhorn / matched-feature phases."""
    t0 = time.perf_counter()
    ngpu = len(features_a)
    half_ngpu = ngpu // 2
    fa_batch1 = torch.cat(features_a[:half_ngpu], 0)                                    # :16-19
    fa_batch2 = torch.cat(features_a[half_ngpu:], 0)
    fb_batch1 = torch.cat(features_b[:half_ngpu], 0)
    fb_batch2 = torch.cat(features_b[half_ngpu:], 0)
    dist_a1_a2, dist_b2_b1, dist_a1_b1, dist_a1_b2, dist_a2_b1, dist_a2_b2 = [], [], [], [], [], []
    for i in range(half_ngpu):                                                          # :29-33
        dist_a1_a2.append(1. - torch.matmul(features_a[i], fa_batch2.t()))
        dist_a1_b1.append(1. - torch.matmul(features_a[i], fb_batch1.t()))
        dist_a1_b2.append(1. - torch.matmul(features_a[i], fb_batch2.t()))
    for i in range(half_ngpu, 2 * half_ngpu):                                           # :35-39
        dist_a2_b1.append(1. - torch.matmul(features_a[i], fb_batch1.t()))
        dist_a2_b2.append(1. - torch.matmul(features_a[i], fb_batch2.t()))
        dist_b2_b1.append(1. - torch.matmul(features_b[i], fb_batch1.t()))
    distances = [torch.cat(dist_a1_a2, 0), torch.cat(dist_b2_b1, 0), torch.cat(dist_a1_b1, 0),
                 torch.cat(dist_a1_b2, 0), torch.cat(dist_a2_b1, 0), torch.cat(dist_a2_b2, 0)]   # :41-43
    t1 = time.perf_counter()
    assignments, entropy = [], []
    for i in range(len(distances)):                                                     # :46-57
        log_a = -sinkhorn_lambda * distances[i]
        for it in range(nr_sinkhorn_iter):
            log_a = log_a - torch.logsumexp(log_a, dim=1, keepdim=True)
            log_a = log_a - torch.logsumexp(log_a, dim=0, keepdim=True)
        assignments.append(torch.softmax(log_a, dim=-1))
        entropy.append(torch.mean(-(assignments[-1] * torch.log_softmax(log_a, dim=-1)).sum(-1)))
    a1a2, b2b1, a1b1, a1b2, a2b1, a2b2 = assignments
    entropy = sum(entropy) / len(entropy)                                               # :61
    t2 = time.perf_counter()
    sp = lambda x: list(torch.chunk(x, half_ngpu, 0))
    f_a1_a2 = sp(torch.matmul(a1a2, fa_batch2))                                         # :64-75
    f_b1_b2 = sp(torch.matmul(b2b1.t(), fb_batch2))
    f_a1_b1 = sp(torch.matmul(a1b1, fb_batch1))
    f_a1_b2 = sp(torch.matmul(a1b2, fb_batch2))
    f_a2_b1 = sp(torch.matmul(a2b1, fb_batch1))
    f_a2_b2 = sp(torch.matmul(a2b2, fb_batch2))
    f_a2_a1 = sp(torch.matmul(a1a2.t(), fa_batch1))
    f_b2_b1 = sp(torch.matmul(b2b1, fb_batch1))
    f_b1_a1 = sp(torch.matmul(a1b1.t(), fa_batch1))
    f_b2_a1 = sp(torch.matmul(a1b2.t(), fa_batch1))
    f_b1_a2 = sp(torch.matmul(a2b1.t(), fa_batch2))
    f_b2_a2 = sp(torch.matmul(a2b2.t(), fa_batch2))
    features_a_a = f_a1_a2 + f_a2_a1                                                    # :78-83
    features_b_b = f_b1_b2 + f_b2_b1
    features_a_b = [0.5 * (f1 + f2) for f1, f2 in zip(f_a1_b1 + f_a2_b1, f_a1_b2 + f_a2_b2)]
    features_b_a = [0.5 * (f1 + f2) for f1, f2 in zip(f_b1_a1 + f_b2_a1, f_b1_a2 + f_b2_a2)]
    if phases is not None:
        phases[:] = [t1 - t0, t2 - t1, time.perf_counter() - t2]
    return features_a_a, features_b_b, features_a_b, features_b_a, entropy


def calc_distance(features_a, features_b, matched_features):
    """utils/matching.py:139-153."""
    ngpu = len(features_a)
    batch_size = features_a[0].shape[0]
    features_a_a, features_b_b, features_a_b, features_b_a, _ = matched_features
    dist = []
    for i in range(ngpu):
        nd_a_a = torch.sum(features_a[i] * features_a_a[i])
        nd_b_b = torch.sum(features_b[i] * features_b_b[i])
        nd_a_b = torch.sum(features_a[i] * features_a_b[i])
        dist.append(nd_b_b + nd_a_a - 2. * nd_a_b)
    return sum(dist) / (2 * batch_size * ngpu)
