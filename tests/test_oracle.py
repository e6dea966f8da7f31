"""CPU tests of the oracle (oracle/): algebraic invariants of the reference algorithm, numpy <-> C cross-check, and the
committed golden vectors.  The reference has no tests of its own, so these are what pins the oracle."""
import glob
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import matching_oracle as mo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _feats(N, D, G, kind="clustered"):
    A = mo.synth_embeddings(N, D, 1, kind)
    B = mo.synth_embeddings(N, D, 2, kind)
    return A, B, list(np.split(A, G)), list(np.split(B, G))


def test_rows_of_plans_sum_to_one_and_entropy_bounds():
    _, _, fa, fb = _feats(32, 64, 4)
    res, plans, _ = mo.get_matched_features(fa, fb, 500.0, 50, np.float64, True)
    for p in plans:
        np.testing.assert_allclose(p.sum(axis=1), 1.0, atol=1e-12)
        assert (p >= 0).all()
    assert 0.0 < res[4] <= np.log(16) + 1e-9


def test_T0_is_row_softmax_of_minus_lambda_C():
    _, _, fa, fb = _feats(16, 32, 2)
    _, plans, dists = mo.get_matched_features(fa, fb, 30.0, 0, np.float64, True)
    for p, c in zip(plans, dists):
        np.testing.assert_allclose(p, mo.softmax(-30.0 * c), atol=1e-14)


def test_pc_identity_equals_calc_distance():
    """calc_distance == (sum_cross <P,C> - 2<P0,C0> - 2<P1,C1>) / 2N."""
    A, _, fa, fb = _feats(24, 48, 4)
    res, plans, dists = mo.get_matched_features(fa, fb, 500.0, 30, np.float64, True)
    d1 = mo.calc_distance(fa, fb, res)
    d2 = mo.distance_from_plans(plans, dists, A.shape[0])
    assert abs(d1 - d2) < 1e-13


def test_fused_grad_equals_difference_of_matched_features():
    """grad_ys written as four 3-term block products."""
    _, _, fa, fb = _feats(16, 40, 2)
    res, plans, _ = mo.get_matched_features(fa, fb, 200.0, 25, np.float64, True)
    ga, gb = mo.grad_features(res)
    fa1, fa2, fb1, fb2, _ = mo.two_batch_blocks(fa, fb)
    Ga, Gb = mo.fused_grad_features(plans, fa1, fa2, fb1, fb2)
    np.testing.assert_allclose(np.concatenate(ga), Ga, atol=1e-14)
    np.testing.assert_allclose(np.concatenate(gb), Gb, atol=1e-14)


def test_tower_split_invariance():
    """The result depends only on the concatenated halves, not on how many towers they are split into."""
    A, B, _, _ = _feats(32, 24, 2)
    r2 = mo.get_matched_features(list(np.split(A, 2)), list(np.split(B, 2)), 100.0, 10)
    r8 = mo.get_matched_features(list(np.split(A, 8)), list(np.split(B, 8)), 100.0, 10)
    for i in range(4):
        np.testing.assert_allclose(np.concatenate(r2[i]), np.concatenate(r8[i]), atol=1e-14)
    assert abs(r2[4] - r8[4]) < 1e-14


def test_permutation_equivariance_within_a_half():
    A, B, _, _ = _feats(16, 32, 2)
    perm = np.random.RandomState(0).permutation(8)
    A2 = A.copy()
    A2[:8] = A[:8][perm]
    r = mo.get_matched_features(list(np.split(A, 2)), list(np.split(B, 2)), 100.0, 10)
    rp = mo.get_matched_features(list(np.split(A2, 2)), list(np.split(B, 2)), 100.0, 10)
    np.testing.assert_allclose(np.concatenate(rp[0])[:8], np.concatenate(r[0])[:8][perm], atol=1e-13)
    np.testing.assert_allclose(np.concatenate(rp[3]), np.concatenate(r[3]), atol=1e-13)     # f_ba sums over a-rows


def test_single_batch_has_no_self_matches():
    _, _, fa, fb = _feats(12, 32, 3)
    _, plans, _ = mo.get_matched_features_single_batch(fa, fb, 500.0, 10, np.float64, True)
    assert np.abs(np.diag(plans[0])).max() == 0.0 and np.abs(np.diag(plans[1])).max() == 0.0
    assert np.abs(np.diag(plans[2])).max() > 0.0


def test_random_matching_rotates_towers():
    fa = [np.full((2, 3), i, np.float32) for i in range(4)]
    fb = [np.full((2, 3), 10 + i, np.float32) for i in range(4)]
    aa, bb, ab, ba, e = mo.get_matched_features_random(fa, fb)
    assert [int(x[0, 0]) for x in aa] == [1, 2, 3, 0] and [int(x[0, 0]) for x in bb] == [11, 12, 13, 10]
    assert ab is fb and ba is fa and e == 0.0


@pytest.mark.parametrize("N,D,lam,T", [(16, 37, 100.0, 7), (64, 512, 500.0, 40), (22, 1000, 500.0, 10)])
def test_c_oracle_matches_numpy_oracle_fp64(N, D, lam, T):
    A, B, fa, fb = _feats(N, D, 2)
    res, plans, dists = mo.get_matched_features(fa, fb, lam, T, np.float64, True)
    c = co.two_batch(A, B, lam, T, dtype=np.float64)
    for i, k in enumerate(["f_aa", "f_bb", "f_ab", "f_ba"]):
        np.testing.assert_allclose(c[k], np.concatenate(res[i]), atol=1e-12)
    np.testing.assert_allclose(c["P"], np.stack(plans), atol=1e-12)
    assert abs(c["entropy"] - res[4]) < 1e-12
    assert abs(c["dist"] - mo.calc_distance(fa, fb, res)) < 1e-12


def test_c_oracle_single_batch_and_euclid_fp64():
    A, B, fa, fb = _feats(12, 48, 3)
    res = mo.get_matched_features_single_batch(fa, fb, 500.0, 15)
    c = co.single_batch(A, B, 500.0, 15, dtype=np.float64)
    for i, k in enumerate(["f_aa", "f_bb", "f_ab", "f_ba"]):
        np.testing.assert_allclose(c[k], np.concatenate(res[i]), atol=1e-12)
    rng = np.random.RandomState(3)
    X, Y = rng.randn(16, 16), rng.randn(16, 16) + 1
    t = mo.toy_get_matched_features(X, Y, 50.0, 10)
    c = co.two_batch(X, Y, 50.0, 10, cost_kind=1, dtype=np.float64)
    for i, k in enumerate(["f_aa", "f_bb", "f_ab", "f_ba"]):
        np.testing.assert_allclose(c[k], t[i], atol=1e-12)


def test_fp32_oracle_is_within_the_parity_gate_of_fp64():
    """The tolerance model of the GPU parity tests must at least admit a correct fp32 CPU run."""
    A, B, fa, fb = _feats(64, 2048, 2, "iid")
    r64, p64, _ = mo.get_matched_features(fa, fb, 500.0, 100, np.float64, True)
    r32, p32, _ = mo.get_matched_features(fa, fb, 500.0, 100, np.float32, True)
    for i in range(4):
        ref = np.concatenate(r64[i])
        assert np.abs(np.concatenate(r32[i]) - ref).max() / np.abs(ref).max() < 5e-5
    assert np.abs(np.stack(p32) - np.stack(p64)).max() / np.stack(p64).max() < 1e-4
    assert abs(mo.calc_distance(fa, fb, r32, np.float32) - mo.calc_distance(fa, fb, r64)) < 1e-6


@pytest.mark.parametrize("path", sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if not os.path.basename(p).startswith("ref_")))
def test_golden_vectors(path):
    g = np.load(path)
    name = os.path.basename(path)
    A, B, lam, T, G = g["A"], g["B"], float(g["lam"]), int(g["T"]), int(g["G"])
    if name.startswith("toy"):
        res = mo.toy_get_matched_features(A, B, lam, T)
        dist = mo.toy_calc_distance(A, B, res)
        cat = lambda x: x
    elif name.startswith("single"):
        fa, fb = list(np.split(A, G)), list(np.split(B, G))
        res = mo.get_matched_features_single_batch(fa, fb, lam, T)
        dist = mo.calc_distance(fa, fb, res)
        cat = np.concatenate
    else:
        fa, fb = list(np.split(A, G)), list(np.split(B, G))
        res = mo.get_matched_features(fa, fb, lam, T)
        dist = mo.calc_distance(fa, fb, res)
        cat = np.concatenate
    for i, k in enumerate(["f_aa", "f_bb", "f_ab", "f_ba"]):
        np.testing.assert_allclose(cat(res[i]), g[k], rtol=0, atol=1e-13)
    assert abs(res[4] - float(g["entropy"])) < 1e-13 and abs(dist - float(g["dist"])) < 1e-13
