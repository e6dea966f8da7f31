"""CPU oracle for the OT-GAN matching hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This file is an op-for-op numpy restatement of the reference's algorithm.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it; the product path (``otgan_b200``) never does.

PARITY PIN: the reference (openai/ot-gan) ships no tests, golden vectors or known-answer values, and TensorFlow 1.x
(un-vendored, un-pinned) is not installable here (python 3.12, no wheel, no network).  The oracle is pinned instead to
the reference's own CODE: tests/golden/make_reference_golden.py imports /root/reference/utils/matching.py,
toy_example/matching_cpu.py, utils/nn.py and models/*.py UNMODIFIED over tests/golden/tf1_numpy_shim.py (a float64 numpy
emulation of the ~35 TensorFlow-1.x primitives they call, semantics documented there) and commits the outputs as
tests/golden/ref_*.npz; tests/test_reference_golden.py checks this oracle against them to 1e-11 (and the CUDA path on
the GPU).  What remains unpinned is only the arithmetic INSIDE the TensorFlow primitives (matmul, logsumexp, softmax,
conv2d), restated in the shim.  Further pins: (i) algebraic invariants of the algorithm (tests/test_oracle.py), (ii) an
independent C restatement (oracle/matching_oracle.c), (iii) fp64 golden vectors of this file (tests/golden/*.npz).

Reference lines followed (paths relative to /root/reference):
  get_matched_features_random          utils/matching.py:3-9
  get_matched_features                 utils/matching.py:11-85
  get_matched_features_single_batch    utils/matching.py:88-136
  calc_distance                        utils/matching.py:139-153
  toy_get_matched_features             toy_example/matching_cpu.py:4-95
  toy_get_matched_features_single_batch toy_example/matching_cpu.py:98-152
  toy_calc_distance                    toy_example/matching_cpu.py:155-164
  grad_features                        train.py:107-128
TensorFlow-1.x op semantics restated: tf.reduce_logsumexp = log(sum(exp(x-max)))+max with
max taken over the reduced axis; tf.nn.softmax over the last axis (max-subtracted);
softmax_cross_entropy_with_logits(labels=p, logits=l) = -sum(p * log_softmax(l), -1).

Every function takes ``dtype`` (np.float64 for the parity target, np.float32 for the
"what a correct fp32 implementation looks like" noise floor and the CPU baseline).
Features are lists of G arrays [bs, D] exactly like the reference's per-tower lists.
"""
import numpy as np

__all__ = [
    "reduce_logsumexp", "softmax", "sinkhorn", "cosine_cost", "euclid_mean_cost",
    "get_matched_features_random", "get_matched_features", "get_matched_features_single_batch",
    "calc_distance", "grad_features", "toy_get_matched_features",
    "toy_get_matched_features_single_batch", "toy_calc_distance", "two_batch_blocks",
    "distance_from_plans", "fused_grad_features",
]


# ----------------------------------------------------------------------------- TF op restatements
def reduce_logsumexp(x, axis):
    """tf.reduce_logsumexp(x, axis, keep_dims=True) (utils/matching.py:53-54)."""
    m = np.max(x, axis=axis, keepdims=True)
    return np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True)) + m


def softmax(x):
    """tf.nn.softmax(x) over the last axis (utils/matching.py:56)."""
    e = np.exp(x - np.max(x, axis=-1, keepdims=True))
    return e / np.sum(e, axis=-1, keepdims=True)


def _log_softmax(x):
    z = x - np.max(x, axis=-1, keepdims=True)
    return z - np.log(np.sum(np.exp(z), axis=-1, keepdims=True))


def sinkhorn(dist, sinkhorn_lambda, nr_sinkhorn_iter, dtype=np.float64):
    """One block of utils/matching.py:50-57.  Returns (assignment P, entropy scalar, final log_a)."""
    log_a = (-dtype(sinkhorn_lambda)) * dist.astype(dtype)
    for _ in range(int(nr_sinkhorn_iter)):
        log_a = log_a - reduce_logsumexp(log_a, axis=1)      # :53
        log_a = log_a - reduce_logsumexp(log_a, axis=0)      # :54
    p = softmax(log_a)                                        # :56
    ent = np.mean(-np.sum(p * _log_softmax(log_a), axis=-1))  # :57
    return p, dtype(ent), log_a


def cosine_cost(x, y):
    """1 - x y^T (utils/matching.py:31-39)."""
    return 1.0 - x @ y.T


def euclid_mean_cost(x, y):
    """0.5*mean(x^2)[:,None] + 0.5*mean(y^2)[None,:] - x y^T / n (toy_example/matching_cpu.py:17-45)."""
    n = x.shape[1]
    return (0.5 * np.mean(np.square(x), axis=1, keepdims=True)
            + 0.5 * np.mean(np.square(y), axis=1).reshape(1, -1)
            - (x @ y.T) / n)


# ----------------------------------------------------------------------------- utils/matching.py
def get_matched_features_random(features_a, features_b):
    """utils/matching.py:3-9: rotate the tower list by one; zero entropy."""
    features_a_a = features_a[1:] + features_a[:1]
    features_b_b = features_b[1:] + features_b[:1]
    return features_a_a, features_b_b, features_b, features_a, np.float32(0.0)


def two_batch_blocks(features_a, features_b, dtype=np.float64, cost=cosine_cost):
    """The six cost blocks in the reference's order (utils/matching.py:16-43).

    Returns (fa1, fa2, fb1, fb2, [C0..C5]) with C0=a1a2, C1=b2b1, C2=a1b1, C3=a1b2, C4=a2b1, C5=a2b2."""
    ngpu = len(features_a)
    half = ngpu // 2
    fa1 = np.concatenate([f.astype(dtype) for f in features_a[:half]], axis=0)
    fa2 = np.concatenate([f.astype(dtype) for f in features_a[half:]], axis=0)
    fb1 = np.concatenate([f.astype(dtype) for f in features_b[:half]], axis=0)
    fb2 = np.concatenate([f.astype(dtype) for f in features_b[half:]], axis=0)
    dists = [cost(fa1, fa2), cost(fb2, fb1), cost(fa1, fb1), cost(fa1, fb2), cost(fa2, fb1), cost(fa2, fb2)]
    return fa1, fa2, fb1, fb2, dists


def _combine_two_batch(assignments, fa1, fa2, fb1, fb2):
    """The twelve products and their regrouping (utils/matching.py:64-83), on full half-batches."""
    p_a1a2, p_b2b1, p_a1b1, p_a1b2, p_a2b1, p_a2b2 = assignments
    a1_a2 = p_a1a2 @ fa2
    b1_b2 = p_b2b1.T @ fb2
    a1_b1 = p_a1b1 @ fb1
    a1_b2 = p_a1b2 @ fb2
    a2_b1 = p_a2b1 @ fb1
    a2_b2 = p_a2b2 @ fb2
    a2_a1 = p_a1a2.T @ fa1
    b2_b1 = p_b2b1 @ fb1
    b1_a1 = p_a1b1.T @ fa1
    b2_a1 = p_a1b2.T @ fa1
    b1_a2 = p_a2b1.T @ fa2
    b2_a2 = p_a2b2.T @ fa2
    f_aa = np.concatenate([a1_a2, a2_a1], axis=0)
    f_bb = np.concatenate([b1_b2, b2_b1], axis=0)
    f_ab = 0.5 * (np.concatenate([a1_b1, a2_b1], axis=0) + np.concatenate([a1_b2, a2_b2], axis=0))
    f_ba = 0.5 * (np.concatenate([b1_a1, b2_a1], axis=0) + np.concatenate([b1_a2, b2_a2], axis=0))
    return f_aa, f_bb, f_ab, f_ba


def get_matched_features(features_a, features_b, sinkhorn_lambda, nr_sinkhorn_iter,
                         dtype=np.float64, return_plans=False):
    """utils/matching.py:11-85.  Lists of G arrays [bs, D] in, four lists of G arrays + entropy out."""
    ngpu = len(features_a)
    fa1, fa2, fb1, fb2, dists = two_batch_blocks(features_a, features_b, dtype)
    plans, ents = [], []
    for d in dists:
        p, e, _ = sinkhorn(d, sinkhorn_lambda, nr_sinkhorn_iter, dtype)
        plans.append(p)
        ents.append(e)
    entropy = dtype(sum(ents) / len(ents))
    f_aa, f_bb, f_ab, f_ba = _combine_two_batch(plans, fa1, fa2, fb1, fb2)
    out = tuple(np.split(f, ngpu, axis=0) for f in (f_aa, f_bb, f_ab, f_ba)) + (entropy,)
    if return_plans:
        return out, plans, dists
    return out


def get_matched_features_single_batch(features_a, features_b, sinkhorn_lambda, nr_sinkhorn_iter,
                                      dtype=np.float64, return_plans=False):
    """utils/matching.py:88-136: three N x N blocks, +999 on the aa/bb diagonals."""
    ngpu = len(features_a)
    bs = features_a[0].shape[0]
    fa = np.concatenate([f.astype(dtype) for f in features_a], axis=0)
    fb = np.concatenate([f.astype(dtype) for f in features_b], axis=0)
    eye = np.eye(ngpu * bs, dtype=dtype)
    dists = [cosine_cost(fa, fa) + 999.0 * eye, cosine_cost(fb, fb) + 999.0 * eye, cosine_cost(fa, fb)]
    plans, ents = [], []
    for d in dists:
        p, e, _ = sinkhorn(d, sinkhorn_lambda, nr_sinkhorn_iter, dtype)
        plans.append(p)
        ents.append(e)
    entropy = dtype(sum(ents) / len(ents))
    p_aa, p_bb, p_ab = plans
    out = (np.split(p_aa @ fa, ngpu, 0), np.split(p_bb @ fb, ngpu, 0),
           np.split(p_ab @ fb, ngpu, 0), np.split(p_ab.T @ fa, ngpu, 0), entropy)
    if return_plans:
        return out, plans, dists
    return out


def calc_distance(features_a, features_b, matched_features, dtype=np.float64):
    """utils/matching.py:139-153."""
    ngpu = len(features_a)
    bs = features_a[0].shape[0]
    f_aa, f_bb, f_ab, _f_ba, _ = matched_features
    dist = []
    for i in range(ngpu):
        nd_a_a = np.sum(features_a[i].astype(dtype) * f_aa[i])
        nd_b_b = np.sum(features_b[i].astype(dtype) * f_bb[i])
        nd_a_b = np.sum(features_a[i].astype(dtype) * f_ab[i])
        dist.append(nd_b_b + nd_a_a - 2.0 * nd_a_b)
    return dtype(sum(dist) / (2 * bs * ngpu))


def grad_features(matched_features):
    """train.py:111,125-126: grad_ys(fake) = f_aa - f_ab, grad_ys(real) = f_bb - f_ba (per tower)."""
    f_aa, f_bb, f_ab, f_ba, _ = matched_features
    return ([x - y for x, y in zip(f_aa, f_ab)], [x - y for x, y in zip(f_bb, f_ba)])


# ----------------------------------------------------------------------------- derived identities
def distance_from_plans(plans, dists, n_total):
    """<P,C> form of calc_distance, valid because every row of every plan sums to one:
    dist = (sum_{k=2..5}<P_k,C_k> - 2<P_0,C_0> - 2<P_1,C_1>) / (2N)."""
    pc = [np.sum(p * c) for p, c in zip(plans, dists)]
    return (pc[2] + pc[3] + pc[4] + pc[5] - 2.0 * pc[0] - 2.0 * pc[1]) / (2.0 * n_total)


def fused_grad_features(plans, fa1, fa2, fb1, fb2):
    """grad_ys as four 3-term block products (what otgan_grad_features_f32 computes)."""
    p0, p1, p2, p3, p4, p5 = plans
    ga1 = p0 @ fa2 - 0.5 * (p2 @ fb1) - 0.5 * (p3 @ fb2)
    ga2 = p0.T @ fa1 - 0.5 * (p4 @ fb1) - 0.5 * (p5 @ fb2)
    gb1 = p1.T @ fb2 - 0.5 * (p2.T @ fa1) - 0.5 * (p4.T @ fa2)
    gb2 = p1 @ fb1 - 0.5 * (p3.T @ fa1) - 0.5 * (p5.T @ fa2)
    return np.concatenate([ga1, ga2], 0), np.concatenate([gb1, gb2], 0)


# ----------------------------------------------------------------------------- toy_example/matching_cpu.py
def toy_get_matched_features(features_a, features_b, sinkhorn_lambda, nr_sinkhorn_iter,
                             dtype=np.float64, return_plans=False):
    """toy_example/matching_cpu.py:4-95: single tensors [2h, D], squared-Euclidean/n cost."""
    fa1, fa2 = np.split(features_a.astype(dtype), 2, axis=0)
    fb1, fb2 = np.split(features_b.astype(dtype), 2, axis=0)
    c = euclid_mean_cost
    dists = [c(fa1, fa2), c(fb2, fb1), c(fa1, fb1), c(fa1, fb2), c(fa2, fb1), c(fa2, fb2)]
    plans, ents = [], []
    for d in dists:
        p, e, _ = sinkhorn(d, sinkhorn_lambda, nr_sinkhorn_iter, dtype)
        plans.append(p)
        ents.append(e)
    entropy = dtype(sum(ents) / len(ents))
    out = _combine_two_batch(plans, fa1, fa2, fb1, fb2) + (entropy,)
    if return_plans:
        return out, plans, dists
    return out


def toy_get_matched_features_single_batch(features_a, features_b, sinkhorn_lambda, nr_sinkhorn_iter,
                                          batch_size, dtype=np.float64):
    """toy_example/matching_cpu.py:98-152 (lists of towers, eye sized by the explicit batch_size arg)."""
    ngpu = len(features_a)
    fa = np.concatenate([f.astype(dtype) for f in features_a], axis=0)
    fb = np.concatenate([f.astype(dtype) for f in features_b], axis=0)
    eye = np.eye(batch_size, dtype=dtype)
    dists = [euclid_mean_cost(fa, fa) + 999.0 * eye, euclid_mean_cost(fb, fb) + 999.0 * eye,
             euclid_mean_cost(fa, fb)]
    plans, ents = [], []
    for d in dists:
        p, e, _ = sinkhorn(d, sinkhorn_lambda, nr_sinkhorn_iter, dtype)
        plans.append(p)
        ents.append(e)
    entropy = dtype(sum(ents) / len(ents))
    p_aa, p_bb, p_ab = plans
    return (np.split(p_aa @ fa, ngpu, 0), np.split(p_bb @ fb, ngpu, 0),
            np.split(p_ab @ fb, ngpu, 0), np.split(p_ab.T @ fa, ngpu, 0), entropy)


def toy_calc_distance(features_a, features_b, matched_features, dtype=np.float64):
    """toy_example/matching_cpu.py:155-164: reduce_mean based, divided by 2."""
    f_aa, f_bb, f_ab, _f_ba, _ = matched_features
    nd_a_a = np.mean(features_a.astype(dtype) * f_aa)
    nd_b_b = np.mean(features_b.astype(dtype) * f_bb)
    nd_a_b = np.mean(features_a.astype(dtype) * f_ab)
    return dtype((nd_b_b + nd_a_a - 2.0 * nd_a_b) / 2.0)


# ----------------------------------------------------------------------------- synthetic inputs
def synth_embeddings(n, d, seed, kind="iid", sigma=0.3, centroid_seed=999, dtype=np.float32):
    """Seeded critic-like embeddings: row-normalised, non-negative (CReLU head, models/dcgan.py:16-19).

    kind="iid": x~N(0,1)[n, d/2]; kind="clustered": 10 centroids (fixed seed) + sigma*noise."""
    rng = np.random.RandomState(seed)
    half = d // 2
    if kind == "iid":
        x = rng.randn(n, half)
    elif kind == "clustered":
        cent = np.random.RandomState(centroid_seed).randn(10, half)
        x = cent[rng.randint(0, 10, size=n)] + sigma * rng.randn(n, half)
    else:
        raise ValueError(kind)
    f = np.concatenate([np.maximum(x, 0.0), np.maximum(-x, 0.0)], axis=1)
    if f.shape[1] < d:  # odd d: pad with a small positive column so rows stay non-zero
        f = np.concatenate([f, np.full((n, d - f.shape[1]), 0.1)], axis=1)
    f = f / np.sqrt(np.sum(np.square(f), axis=1, keepdims=True))
    return f.astype(dtype)
