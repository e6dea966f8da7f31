"""Property tests (hypothesis, CPU): the invariants the matching oracle is pinned by, over RANDOM shapes / lambda / T rather
than the handful of fixed cases of tests/test_oracle.py ("property" level).  The reference ships no tests
and cannot run here, so these algebraic properties -- rows of every plan sum to one, T = 0 is a row softmax, the <P,C> form
equals calc_distance, the fused gradient equals f_aa - f_ab / f_bb - f_ba, tower-split invariance, numpy == C restatement --
are what stands between the oracle and a silent mistake (utils/matching.py:11-153, train.py:111,125-126)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import c_oracle as co
from oracle import matching_oracle as mo

shapes = st.tuples(st.integers(1, 6),            # rows per half-tower chunk
                   st.sampled_from([2, 4]),      # towers G (even, utils/matching.py:12-13)
                   st.integers(2, 24),           # feature width D
                   st.floats(1.0, 300.0),        # sinkhorn_lambda
                   st.integers(0, 12),           # nr_sinkhorn_iter
                   st.integers(0, 2 ** 16))      # seed


def _features(bs, G, D, seed):
    rng = np.random.RandomState(seed)
    def unit(n):
        x = np.abs(rng.randn(n, D)) + 1e-3          # non-negative like the CReLU head, then L2-normalised
        return x / np.linalg.norm(x, axis=1, keepdims=True)
    a, b = unit(bs * G), unit(bs * G)
    return list(np.split(a, G)), list(np.split(b, G))


@settings(max_examples=40, deadline=None)
@given(shapes)
def test_plan_and_distance_invariants(case):
    bs, G, D, lam, T, seed = case
    fa, fb = _features(bs, G, D, seed)
    out, plans, dists = mo.get_matched_features(fa, fb, lam, T, np.float64, return_plans=True)
    h = bs * G // 2
    for p, c in zip(plans, dists):
        assert p.shape == (h, h) and np.all(p >= 0)
        np.testing.assert_allclose(p.sum(1), 1.0, rtol=0, atol=1e-12)                  # trailing ROW softmax (:56)
        if T == 0:
            np.testing.assert_allclose(p, mo.softmax(-lam * c), rtol=0, atol=1e-14)
    assert 0.0 <= out[4] <= np.log(h) + 1e-12                                           # entropy of h-way distributions
    # <P,C> identity == calc_distance (App. A.3), fused gradient == difference of matched features (App. A.4)
    d_ref = mo.calc_distance(fa, fb, out)
    assert abs(mo.distance_from_plans(plans, dists, bs * G) - d_ref) < 1e-12
    fa1, fa2, fb1, fb2, _ = mo.two_batch_blocks(fa, fb)
    ga, gb = mo.fused_grad_features(plans, fa1, fa2, fb1, fb2)
    ra, rb = mo.grad_features(out)
    np.testing.assert_allclose(ga, np.concatenate(ra), rtol=0, atol=1e-13)
    np.testing.assert_allclose(gb, np.concatenate(rb), rtol=0, atol=1e-13)


@settings(max_examples=25, deadline=None)
@given(shapes)
def test_tower_split_invariance_and_c_restatement(case):
    bs, G, D, lam, T, seed = case
    fa, fb = _features(bs, G, D, seed)
    out = mo.get_matched_features(fa, fb, lam, T, np.float64)
    # the same rows presented as 2 towers give the same matched features (the list structure only fixes the halves)
    A, B = np.concatenate(fa), np.concatenate(fb)
    out2 = mo.get_matched_features(list(np.split(A, 2)), list(np.split(B, 2)), lam, T, np.float64)
    for i in range(4):
        np.testing.assert_allclose(np.concatenate(out[i]), np.concatenate(out2[i]), rtol=0, atol=1e-13)
    # independent C + OpenMP restatement (float64)
    r = co.two_batch(A, B, lam, T, dtype=np.float64)
    np.testing.assert_allclose(r["f_aa"], np.concatenate(out[0]), rtol=0, atol=1e-10)
    np.testing.assert_allclose(r["f_ab"], np.concatenate(out[2]), rtol=0, atol=1e-10)
    assert abs(r["entropy"] - out[4]) < 1e-10 and abs(r["dist"] - mo.calc_distance(fa, fb, out)) < 1e-10
