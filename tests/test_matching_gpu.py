"""GPU parity tests (-m gpu): every CUDA kernel of the matching path, called through the C ABI (ctypes), against the
fp64 oracle on identical seeded inputs, plus the committed golden vectors and size-independent properties at the
BASELINE.json sizes.

Tolerance model (everything is max-abs-error / max-abs-value against the fp64 oracle):
  cost blocks C          <= 2e-6        (the fp32 oracle sits at ~1.2e-7 .. 1e-6 depending on D)
  plans P                <= 1e-4        (lambda = 500 amplifies the fp32 rounding of C 500x; fp32 oracle: 1.5e-5 .. 6e-5)
  matched features/grads <= max(1e-5, 1.5 x the fp32 oracle's own error on the same inputs)   (SURVEY 8d gate; the fp32
                            oracle sits at 2e-6 .. 2e-5 -- lambda = 500 amplifies fp32 rounding of C inside log P);
                            kernel-level plan-apply tests use fixed 3e-6 / 1e-5 gates
  entropy                <= 5e-6 relative
  distance               <= 1e-6 ABSOLUTE (it is a cancellation of O(1) terms down to ~1e-4)
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import matching_oracle as mo

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL_C, TOL_P, TOL_F, TOL_ENT, TOL_DIST = 2e-6, 1e-4, 1e-5, 5e-6, 1e-6


def tol_f(ref64, *ref32s):
    """SURVEY 8d parity gate for matched features / grad_ys: max(1e-5, 1.5 x the fp32 oracle's own error vs fp64).  Where
    several fp32 oracles exist (numpy, the independent C restatement, the torch-CPU op-for-op restatement) the noise floor
    is the largest of their errors: correct fp32 implementations of this path differ from each other by that much
    (lambda = 500 puts |log_a| at ~350, where one fp32 ulp is 3e-5)."""
    worst = 0.0
    for ref32 in ref32s:
        for a, b in zip(ref64, ref32):
            a, b = np.concatenate(a) if isinstance(a, list) else a, np.concatenate(b) if isinstance(b, list) else b
            worst = max(worst, float(np.abs(np.asarray(b, dtype=np.float64) - a).max() / np.abs(a).max()))
    return max(TOL_F, 1.5 * worst)


def torch_fp32_two_batch(fa, fb, lam, T):
    """The torch-CPU op-for-op restatement's fp32 matched features (third fp32 noise-floor sample)."""
    from oracle import torch_oracle as to
    r = to.get_matched_features([torch.from_numpy(x) for x in fa], [torch.from_numpy(x) for x in fb], lam, T)
    return [torch.cat(r[i]).numpy() for i in range(4)]


def c_fp32_two_batch(A, B, lam, T):
    """The C restatement's fp32 matched features (second fp32 noise-floor sample)."""
    from oracle import c_oracle as co
    r = co.two_batch(A, B, lam, T, want_plans=False)
    return [r["f_aa"], r["f_bb"], r["f_ab"], r["f_ba"]]


def relerr(a, ref):
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-300))


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def towers(A, G):
    return list(torch.chunk(dev(A), G, dim=0))


@pytest.fixture(scope="module")
def M():
    from otgan_b200.utils import matching
    return matching


# ----------------------------------------------------------------------------------------------- kernel level
@pytest.mark.parametrize("rows,cols,D,kind", [(16, 16, 64, 0), (12, 20, 37, 0), (128, 128, 4096, 0), (64, 64, 32768, 0),
                                              (32, 32, 16, 1), (24, 24, 100, 1), (130, 70, 260, 0)])
def test_cost_blocks_kernel(M, rows, cols, D, kind):
    from otgan_b200 import _lib
    rng = np.random.RandomState(rows + D)
    if kind == 0:
        X = [mo.synth_embeddings(rows, D, 10 + k, "iid") for k in range(3)]
        Y = [mo.synth_embeddings(cols, D, 20 + k, "iid") for k in range(3)]
    else:
        X = [rng.randn(rows, D).astype(np.float32) for _ in range(3)]
        Y = [(rng.randn(cols, D) + 0.5).astype(np.float32) for _ in range(3)]
    lam = 50.0
    L = M.cost_blocks([dev(x) for x in X], [dev(y) for y in Y], lam, kind, None, _lib.IMPL_SIMT)
    torch.cuda.synchronize()
    cost = mo.cosine_cost if kind == 0 else mo.euclid_mean_cost
    for k in range(3):
        C = cost(X[k].astype(np.float64), Y[k].astype(np.float64))
        got = L[k].cpu().double().numpy() / -lam
        assert np.abs(got - C).max() / np.abs(C).max() < TOL_C


@pytest.mark.parametrize("h,D", [(128, 32768), (128, 7296), (64, 32768), (128, 4096), (96, 2052), (256, 8192), (200, 2304), (128, 131072)])
def test_cost_blocks_tcgen05_unit_rows_sweep(M, h, D):
    """The 3xTF32 tensor-core cost kernel on L2-normalised rows (what the critic head emits) in the six-block two-batch pattern:
    ragged tiles (h = 96, 200), K tails that are not a multiple of the 32-element chunk (D = 2052), cfg4's D = 7296 and cfg5's
    D = 131072 against the fp64 oracle.  Also the single-batch X == Y case (diagonal tiles load one operand, +999 diagonal)."""
    from otgan_b200 import _lib
    A, B = mo.synth_embeddings(2 * h, D, 1, "clustered", sigma=1.0), mo.synth_embeddings(2 * h, D, 2, "clustered", sigma=1.0)
    Ad, Bd = dev(A), dev(B)
    a1, a2, b1, b2 = Ad[:h], Ad[h:], Bd[:h], Bd[h:]
    X, Y = [a1, b2, a1, a1, a2, a2], [a2, b1, b1, b2, b1, b2]
    lam = 500.0
    Lh = M.cost_blocks(X, Y, lam, 0, None, _lib.IMPL_TCGEN05)
    torch.cuda.synchronize()
    A64, B64 = A.astype(np.float64), B.astype(np.float64)
    xs = [A64[:h], B64[h:], A64[:h], A64[:h], A64[h:], A64[h:]]
    ys = [A64[h:], B64[:h], B64[:h], B64[h:], B64[:h], B64[h:]]
    for k in range(6):
        assert relerr(Lh[k] / -lam, mo.cosine_cost(xs[k], ys[k])) < TOL_C, ("tcgen05 vs fp64", k)
    if h <= 128 and D <= 8192:
        Ls = M.cost_blocks([Ad, Bd, Ad], [Ad, Bd, Bd], lam, 0, [999.0, 999.0, 0.0], _lib.IMPL_TCGEN05)
        torch.cuda.synchronize()
        for k, (x, y, dg) in enumerate(((A64, A64, 999.0), (B64, B64, 999.0), (A64, B64, 0.0))):
            C = mo.cosine_cost(x, y) + dg * np.eye(2 * h)
            assert relerr(Ls[k] / -lam, C) < TOL_C, ("tcgen05 single-batch vs fp64", k)


@pytest.mark.parametrize("h,D,kind", [(128, 32768, 0), (128, 7296, 0), (64, 32768, 0), (128, 4096, 0), (96, 1000, 0),
                                      (128, 48, 0), (32, 64, 1), (100, 520, 1), (256, 8192, 0), (200, 1000, 0), (384, 2048, 1)])
def test_cost_blocks_tcgen05_two_batch(M, h, D, kind):
    """TMA + tcgen05 3xTF32 cost kernel on the six-block two-batch pattern: against the fp64 oracle and against the
    exact-fp32 SIMT kernel (the two must agree to fp32 noise; lambda = 500 then keeps P within its gate)."""
    from otgan_b200 import _lib
    if kind == 0:
        A, B = mo.synth_embeddings(2 * h, D, 1, "clustered", sigma=1.0), mo.synth_embeddings(2 * h, D, 2, "clustered", sigma=1.0)
    else:
        rng = np.random.RandomState(h + D)
        A, B = rng.randn(2 * h, D).astype(np.float32), (rng.randn(2 * h, D) * 0.5 + 1).astype(np.float32)
    Ad, Bd = dev(A), dev(B)
    a1, a2, b1, b2 = Ad[:h], Ad[h:], Bd[:h], Bd[h:]
    X, Y = [a1, b2, a1, a1, a2, a2], [a2, b1, b1, b2, b1, b2]
    lam = 500.0 if kind == 0 else 50.0
    Ltc = M.cost_blocks(X, Y, lam, kind, None, _lib.IMPL_TCGEN05).clone()
    Lsi = M.cost_blocks(X, Y, lam, kind, None, _lib.IMPL_SIMT).clone()
    torch.cuda.synchronize()
    cost = mo.cosine_cost if kind == 0 else mo.euclid_mean_cost
    A64, B64 = A.astype(np.float64), B.astype(np.float64)
    xs = [A64[:h], B64[h:], A64[:h], A64[:h], A64[h:], A64[h:]]
    ys = [A64[h:], B64[:h], B64[:h], B64[h:], B64[:h], B64[h:]]
    for k in range(6):
        C = cost(xs[k], ys[k])
        # unnormalised Gaussian rows (the toy Euclidean cost) at h = 384, D = 2048 sit right at 2e-6 for BOTH kernels (fp32
        # accumulation of 2048 O(1) products); the cosine blocks of the real path stay below TOL_C
        tol_c = TOL_C if kind == 0 else 1.5 * TOL_C
        assert relerr(Ltc[k] / -lam, C) < tol_c, ("tcgen05 vs fp64", k)
        assert relerr(Lsi[k] / -lam, C) < tol_c, ("simt vs fp64", k)
    assert float((Ltc - Lsi).abs().max()) / lam < 1e-6


def test_cost_blocks_tcgen05_single_batch_pattern(M):
    from otgan_b200 import _lib
    A, B = mo.synth_embeddings(96, 2048, 1, "clustered"), mo.synth_embeddings(96, 2048, 2, "clustered")
    Ad, Bd = dev(A), dev(B)
    L = M.cost_blocks([Ad, Bd, Ad], [Ad, Bd, Bd], 500.0, 0, [999.0, 999.0, 0.0], _lib.IMPL_TCGEN05)
    a, b = A.astype(np.float64), B.astype(np.float64)
    refs = [mo.cosine_cost(a, a) + 999 * np.eye(96), mo.cosine_cost(b, b) + 999 * np.eye(96), mo.cosine_cost(a, b)]
    for k in range(3):
        assert relerr(L[k] / -500.0, refs[k]) < TOL_C


def test_cost_blocks_diag_and_strided_rows(M):
    from otgan_b200 import _lib
    Z = dev(mo.synth_embeddings(40, 96, 3, "iid"))
    A, B = Z[:20, :64], Z[20:, :64]            # row stride 96 != D, base pointers still 16B aligned
    L = M.cost_blocks([A, B], [A, B], 500.0, 0, [999.0, 0.0], _lib.IMPL_SIMT)
    a, b = A.cpu().double().numpy(), B.cpu().double().numpy()
    C0 = mo.cosine_cost(a, a) + 999.0 * np.eye(20)
    C1 = mo.cosine_cost(b, b)
    assert relerr(L[0] / -500.0, C0) < TOL_C and relerr(L[1] / -500.0, C1) < TOL_C


@pytest.mark.parametrize("impl", [0, 1])      # 0 = AUTO (scaling-form kernel), 1 = SIMT (literal log-domain kernel)
@pytest.mark.parametrize("nblk,rows,cols,lam,T", [(6, 128, 128, 500.0, 100), (6, 64, 64, 500.0, 100), (3, 32, 32, 50.0, 10),
                                                  (2, 17, 23, 100.0, 5), (1, 128, 128, 500.0, 0), (6, 125, 125, 500.0, 30),
                                                  (2, 128, 128, 500.0, 1), (2, 100, 128, 2000.0, 50)])
def test_sinkhorn_kernel(M, nblk, rows, cols, lam, T, impl):
    D = 256
    C = np.stack([mo.cosine_cost(mo.synth_embeddings(rows, D, 30 + k, "clustered", sigma=1.0).astype(np.float64),
                                 mo.synth_embeddings(cols, D, 40 + k, "clustered", sigma=1.0).astype(np.float64))
                  for k in range(nblk)])
    L0 = dev((-lam * C).astype(np.float32))
    P, ent, pc = M.sinkhorn(L0, lam, T, True, impl)
    torch.cuda.synchronize()
    C32 = L0.cpu().double().numpy() / -lam        # the oracle sees exactly the fp32 cost the kernel saw
    for k in range(nblk):
        p, e, _ = mo.sinkhorn(C32[k], lam, T, np.float64)
        scale = max(1.0, lam / 500.0)           # fp32 rounding of -lambda*C grows with lambda
        assert relerr(P[k], p) < TOL_P * scale
        assert abs(float(ent[k]) - e) <= TOL_ENT * scale * max(abs(e), 1e-3)
        assert abs(float(pc[k]) - np.sum(p * C32[k])) < 2e-5 * rows


def test_sinkhorn_single_batch_diagonal_and_hard_inputs(M):
    """+999 on the diagonal (-lambda*C = -499500 there), near-duplicate points, and a very peaked problem."""
    rng = np.random.RandomState(11)
    X = mo.synth_embeddings(96, 512, 5, "clustered", sigma=0.05).astype(np.float64)     # tight clusters: near-ties
    C0 = mo.cosine_cost(X, X) + 999.0 * np.eye(96)
    C1 = rng.rand(96, 96) * 2.0                                                          # spread costs: peaked plans
    L0 = dev(np.stack([-500.0 * C0, -500.0 * C1]).astype(np.float32))
    for impl in (0, 1):
        P, ent, _ = M.sinkhorn(L0, 500.0, 60, True, impl)
        C32 = L0.cpu().double().numpy() / -500.0
        for k in range(2):
            p, e, _ = mo.sinkhorn(C32[k], 500.0, 60, np.float64)
            assert relerr(P[k], p) < TOL_P
            assert abs(float(ent[k]) - e) <= 2e-5 * max(abs(e), 1e-3) + 1e-6
        assert float(torch.diagonal(P[0]).abs().max()) == 0.0


def test_sinkhorn_fast_path_is_taken(M):
    """The scaling-form kernel may fall back to log-domain half-steps only while the potentials are still moving."""
    C = np.stack([mo.cosine_cost(mo.synth_embeddings(128, 1024, 50 + k, "clustered", sigma=1.0).astype(np.float64),
                                 mo.synth_embeddings(128, 1024, 60 + k, "clustered", sigma=1.0).astype(np.float64))
                  for k in range(6)])
    L0 = dev((-500.0 * C).astype(np.float32))
    _, _, _, slow = M.sinkhorn(L0, 500.0, 100, True, 0, want_stats=True)
    slow = slow.cpu().numpy()
    assert (slow >= 1).all() and (slow <= 40).all(), slow           # 200 half-steps in total


@pytest.mark.parametrize("impl", [0, 1])      # 0 = AUTO (8-CTA cluster kernel up to side 512), 1 = SIMT (streaming kernels)
@pytest.mark.parametrize("nblk,rows,cols,lam,T", [(3, 256, 256, 500.0, 100), (2, 200, 150, 500.0, 20), (1, 1024, 1024, 500.0, 10),
                                                  (2, 129, 300, 100.0, 5), (1, 256, 256, 500.0, 0), (6, 256, 256, 500.0, 100),
                                                  (2, 512, 512, 500.0, 100), (2, 100, 130, 500.0, 30), (1, 511, 257, 500.0, 7),
                                                  (3, 130, 129, 50.0, 1)])
def test_sinkhorn_large_blocks(M, nblk, rows, cols, lam, T, impl):
    """Blocks larger than one SM (single-batch N = 256, 64x64-image configs): the persistent cluster kernel (one 8-CTA cluster per
    block, slabs of rows in registers, column partials through distributed shared memory) and the streaming log-domain kernels."""
    D = 128
    C = np.stack([mo.cosine_cost(mo.synth_embeddings(rows, D, 70 + k, "clustered", sigma=1.0).astype(np.float64),
                                 mo.synth_embeddings(cols, D, 80 + k, "clustered", sigma=1.0).astype(np.float64))
                  for k in range(nblk)])
    L0 = dev((-lam * C).astype(np.float32))
    P, ent, pc = M.sinkhorn(L0, lam, T, True, impl)
    torch.cuda.synchronize()
    C32 = L0.cpu().double().numpy() / -lam
    for k in range(nblk):
        p, e, _ = mo.sinkhorn(C32[k], lam, T, np.float64)
        assert relerr(P[k], p) < TOL_P
        assert abs(float(ent[k]) - e) <= TOL_ENT * max(abs(e), 1e-3)
        assert abs(float(pc[k]) - np.sum(p * C32[k])) < 2e-5 * rows
    if impl == 0 and max(rows, cols) <= 512:
        # the cluster kernel keeps the block in registers: the plan is optional, the statistics do not depend on it
        P2, ent2, pc2 = M.sinkhorn(L0, lam, T, False, impl)
        assert P2 is None and torch.equal(ent2, ent) and torch.equal(pc2, pc)


def test_single_batch_at_headline_size(M):
    """utils/matching.py:88-136 at N = 256: three 256 x 256 blocks with +999 on the aa/bb diagonals."""
    N, D, G = 256, 32768, 4
    A, B = mo.synth_embeddings(N, D, 1, "clustered", sigma=1.0), mo.synth_embeddings(N, D, 2, "clustered", sigma=1.0)
    fa, fb = list(np.split(A, G)), list(np.split(B, G))
    ref = mo.get_matched_features_single_batch(fa, fb, 500.0, 100)
    tol = tol_f(ref[:4], mo.get_matched_features_single_batch(fa, fb, 500.0, 100, np.float32)[:4])
    got = M.get_matched_features_single_batch(towers(A, G), towers(B, G), 500.0, 100)
    for i in range(4):
        assert relerr(torch.cat(got[i]), np.concatenate(ref[i])) < tol, (i, tol)
    assert abs(float(got[4]) - ref[4]) <= TOL_ENT * abs(ref[4])


@pytest.mark.parametrize("D", [4096, 131072])
def test_two_batch_h256(M, D):
    """BASELINE config 5 (N = 512, h = 256, D = 131072 = the 64 x 64 critic's feature count, 8 towers) through the public API."""
    N, G = 512, 8
    A, B = mo.synth_embeddings(N, D, 1, "clustered", sigma=1.0), mo.synth_embeddings(N, D, 2, "clustered", sigma=1.0)
    fa, fb = list(np.split(A, G)), list(np.split(B, G))
    ref = mo.get_matched_features(fa, fb, 500.0, 100)
    tol = tol_f(ref[:4], mo.get_matched_features(fa, fb, 500.0, 100, np.float32)[:4])
    ta, tb = towers(A, G), towers(B, G)
    got = M.get_matched_features(ta, tb, 500.0, 100)
    for i in range(4):
        assert relerr(torch.cat(got[i]), np.concatenate(ref[i])) < tol, (i, tol)
    assert abs(float(M.calc_distance(ta, tb, got)) - mo.calc_distance(fa, fb, ref)) < TOL_DIST


def test_sinkhorn_rows_sum_to_one_at_full_size(M):
    rng = np.random.RandomState(0)
    L0 = dev((-500.0 * rng.rand(6, 128, 128)).astype(np.float32))
    P, ent, _ = M.sinkhorn(L0, 500.0, 500)
    s = P.sum(dim=2)
    assert float((s - 1).abs().max()) < 1e-5
    assert float(P.min()) >= 0.0 and bool(torch.isfinite(P).all())
    assert 0.0 <= float(ent.min()) and float(ent.max()) <= np.log(128) + 1e-5


@pytest.mark.parametrize("impl", [1, 2])          # 1 = SIMT (exact fp32), 2 = TMA + tcgen05 3xTF32
@pytest.mark.parametrize("h,D", [(16, 64), (128, 4096), (20, 37), (64, 1000), (128, 32768), (100, 7296), (4, 32), (128, 160),
                                 (256, 4096), (200, 520), (384, 1024)])
def test_plan_apply_kernels(M, h, D, impl):
    from otgan_b200 import _lib
    lib = _lib.load()
    if impl == 2 and D % 4 != 0:
        pytest.skip("tcgen05 path needs 16-byte aligned rows")
    ws, ws_bytes = M._plan_ws(torch.device("cuda", 0), h)
    rng = np.random.RandomState(h * 7 + D)
    P = rng.rand(6, h, h)
    P /= P.sum(axis=2, keepdims=True)
    A, B = rng.randn(2 * h, D), rng.randn(2 * h, D)
    Pd, Ad, Bd = dev(P.astype(np.float32)), dev(A.astype(np.float32)), dev(B.astype(np.float32))
    P64, A64, B64 = Pd.cpu().double().numpy(), Ad.cpu().double().numpy(), Bd.cpu().double().numpy()
    outs = [torch.empty(2 * h, D, device="cuda") for _ in range(4)]
    s = torch.cuda.current_stream().cuda_stream
    rc = lib.otgan_matched_two_batch_f32(h, D, Pd.data_ptr(), Ad.data_ptr(), Bd.data_ptr(), D, outs[0].data_ptr(),
                                         outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(), D, ws.data_ptr(),
                                         ws_bytes, impl, s)
    assert rc == 0, lib.otgan_last_error()
    # SIMT = exact fp32 FMA chains; tcgen05 = 3xTF32 operands (2^-22) + a truncating tensor-core accumulator
    # signed Gaussian features cancel in P*F (|out| ~ 1/sqrt(h)), so the error relative to max|out| grows like sqrt(h) beyond 128
    tol = (3e-6 if impl == 1 else 1e-5) * max(1.0, (h / 128.0) ** 0.5)
    ref = mo._combine_two_batch(list(P64), A64[:h], A64[h:], B64[:h], B64[h:])
    for o, r in zip(outs, ref):
        assert relerr(o, r) < tol
    Ga, Gb = torch.empty(2 * h, D, device="cuda"), torch.empty(2 * h, D, device="cuda")
    rc = lib.otgan_grad_features_f32(h, D, Pd.data_ptr(), Ad.data_ptr(), Bd.data_ptr(), D, Ga.data_ptr(), Gb.data_ptr(), D,
                                     ws.data_ptr(), ws_bytes, impl, s)
    assert rc == 0, lib.otgan_last_error()
    ra, rb = mo.fused_grad_features(list(P64), A64[:h], A64[h:], B64[:h], B64[h:])
    assert relerr(Ga, ra) < tol and relerr(Gb, rb) < tol


@pytest.mark.parametrize("h,D", [(128, 32768), (128, 7296), (64, 4096), (96, 2052), (256, 8192), (200, 2304)])
def test_plan_apply_tcgen05_plan_like_inputs(M, h, D):
    """The tensor-core plan-apply kernel on the inputs the train loop gives it (L2-normalised non-negative feature rows, peaked
    plans): matched features and the fused grad_ys against fp64; ragged h and D tails; and the row-range form (one rank's
    towers) against the full result, bit for bit."""
    from otgan_b200 import _lib
    lib = _lib.load()
    ws, ws_bytes = M._plan_ws(torch.device("cuda", 0), h)
    rng = np.random.RandomState(h * 7 + D)
    P = rng.rand(6, h, h) ** 8                      # peaked rows like a Sinkhorn plan, entries in [0, 1]
    P /= P.sum(axis=2, keepdims=True)
    A, B = mo.synth_embeddings(2 * h, D, 3, "clustered", sigma=1.0), mo.synth_embeddings(2 * h, D, 4, "clustered", sigma=1.0)
    Pd, Ad, Bd = dev(P.astype(np.float32)), dev(A), dev(B)
    P64, A64, B64 = Pd.cpu().double().numpy(), A.astype(np.float64), B.astype(np.float64)
    outs = [torch.empty(2 * h, D, device="cuda") for _ in range(4)]
    s = torch.cuda.current_stream().cuda_stream
    rc = lib.otgan_matched_two_batch_f32(h, D, Pd.data_ptr(), Ad.data_ptr(), Bd.data_ptr(), D, outs[0].data_ptr(),
                                         outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(), D, ws.data_ptr(),
                                         ws_bytes, _lib.IMPL_TCGEN05, s)
    assert rc == 0, lib.otgan_last_error()
    ref = mo._combine_two_batch(list(P64), A64[:h], A64[h:], B64[:h], B64[h:])
    # all-positive products and an accumulator that truncates (cost_tc.cu, note 2): the chain of 3 terms x h / 16 MMAs per
    # output leaves a bias that grows with h (measured 3.5e-6 at h = 256); the features' gate is 1e-5
    tol = 3e-6 * max(1.0, h / 128.0)
    for o, r in zip(outs, ref):
        assert relerr(o, r) < tol
    Ga, Gb = torch.empty(2 * h, D, device="cuda"), torch.empty(2 * h, D, device="cuda")
    rc = lib.otgan_grad_features_f32(h, D, Pd.data_ptr(), Ad.data_ptr(), Bd.data_ptr(), D, Ga.data_ptr(), Gb.data_ptr(), D,
                                     ws.data_ptr(), ws_bytes, _lib.IMPL_TCGEN05, s)
    assert rc == 0, lib.otgan_last_error()
    ra, rb = mo.fused_grad_features(list(P64), A64[:h], A64[h:], B64[:h], B64[h:])
    scale = max(np.abs(ref[0]).max(), np.abs(ref[2]).max())          # grad_ys = f_aa - f_ab: gate relative to the features' scale
    assert np.abs(Ga.cpu().double().numpy() - ra).max() / scale < tol and np.abs(Gb.cpu().double().numpy() - rb).max() / scale < tol
    if h % 2 == 0:
        lo, hi = h // 2, h + h // 2
        Ga2, Gb2 = torch.zeros_like(Ga), torch.zeros_like(Gb)
        rc = lib.otgan_grad_features_rows_f32(h, D, Pd.data_ptr(), Ad.data_ptr(), Bd.data_ptr(), D, Ga2.data_ptr(), Gb2.data_ptr(), D,
                                              lo, hi, ws.data_ptr(), ws_bytes, _lib.IMPL_TCGEN05, s)
        assert rc == 0, lib.otgan_last_error()
        assert torch.equal(Ga2[lo:hi], Ga[lo:hi]) and torch.equal(Gb2[lo:hi], Gb[lo:hi])


def test_matching_step_headline_vs_oracle(M):
    """matching_step (what the train loop calls: fused grad_ys + distance / entropy from <P,C>) against the fp64 oracle at the
    headline size."""
    N, D, G, lam, T = 256, 32768, 2, 500.0, 100
    A, B = mo.synth_embeddings(N, D, 1, "clustered", sigma=1.0), mo.synth_embeddings(N, D, 2, "clustered", sigma=1.0)
    fa, fb = list(np.split(A, G)), list(np.split(B, G))
    ref = mo.get_matched_features(fa, fb, lam, T)
    rga, rgb = mo.grad_features(ref)
    ref32 = mo.get_matched_features(fa, fb, lam, T, dtype=np.float32)
    tol = tol_f(ref[:4], ref32[:4], c_fp32_two_batch(A, B, lam, T))
    scale = np.abs(np.concatenate(ref[0])).max()
    ga, gb, stats = M.matching_step(towers(A, G), towers(B, G), lam, T)
    torch.cuda.synchronize()
    ea = np.abs(torch.cat(ga).cpu().double().numpy() - np.concatenate(rga)).max() / scale
    eb = np.abs(torch.cat(gb).cpu().double().numpy() - np.concatenate(rgb)).max() / scale
    assert ea < tol and eb < tol, (ea, eb, tol)
    assert abs(float(stats[0]) - mo.calc_distance(fa, fb, ref)) < TOL_DIST
    assert abs(float(stats[1]) - ref[4]) < TOL_ENT * abs(ref[4])


# ----------------------------------------------------------------------------------------------- API level
CONFIGS = [
    # N, D, G, lam, T, embeddings           (BASELINE.json configs 2,3(short),4 + headline + small/ragged)
    (128, 32768, 2, 500.0, 100, "iid"),
    (256, 32768, 2, 500.0, 100, "iid"),
    (256, 32768, 8, 500.0, 100, "clustered"),
    (256, 7296, 4, 500.0, 100, "clustered"),
    (256, 32768, 2, 500.0, 500, "clustered"),       # BASELINE config 3 at its real size: T = 500
    (256, 32768, 2, 500.0, 500, "iid"),
    (64, 512, 4, 500.0, 500, "clustered"),
    (12, 37, 2, 100.0, 7, "iid"),
]


@pytest.mark.parametrize("N,D,G,lam,T,kind", CONFIGS)
def test_get_matched_features_vs_fp64_oracle(M, N, D, G, lam, T, kind):
    A = mo.synth_embeddings(N, D, 1, kind, sigma=1.0)
    B = mo.synth_embeddings(N, D, 2, kind, sigma=1.0)
    fa, fb = list(np.split(A, G)), list(np.split(B, G))
    ref = mo.get_matched_features(fa, fb, lam, T)
    ref_dist = mo.calc_distance(fa, fb, ref)
    tol = tol_f(ref[:4], mo.get_matched_features(fa, fb, lam, T, np.float32)[:4], c_fp32_two_batch(A, B, lam, T),
                torch_fp32_two_batch(fa, fb, lam, T))
    ta, tb = towers(A, G), towers(B, G)
    got = M.get_matched_features(ta, tb, lam, T)
    assert len(got) == 5 and all(len(got[i]) == G and got[i][0].shape == (N // G, D) for i in range(4))
    for i in range(4):
        assert relerr(torch.cat(got[i]), np.concatenate(ref[i])) < tol, (i, tol)
    assert abs(float(got[4]) - ref[4]) <= TOL_ENT * abs(ref[4])
    dist = M.calc_distance(ta, tb, got)
    assert abs(float(dist) - ref_dist) < TOL_DIST
    # fused train-loop form: grad_ys + [distance, entropy]
    ga, gb, stats = M.matching_step(ta, tb, lam, T)
    rga, rgb = mo.grad_features(ref)
    scale = max(np.abs(np.concatenate(ref[0])).max(), np.abs(np.concatenate(ref[1])).max())
    assert np.abs(torch.cat(ga).cpu().double().numpy() - np.concatenate(rga)).max() / scale < tol
    assert np.abs(torch.cat(gb).cpu().double().numpy() - np.concatenate(rgb)).max() / scale < tol
    assert abs(float(stats[0]) - ref_dist) < TOL_DIST
    assert abs(float(stats[1]) - ref[4]) <= TOL_ENT * abs(ref[4])


def test_results_are_bitwise_deterministic(M):
    A, B = mo.synth_embeddings(256, 8192, 1), mo.synth_embeddings(256, 8192, 2)
    ta, tb = towers(A, 2), towers(B, 2)
    r1 = M.get_matched_features(ta, tb, 500.0, 50)
    c1 = [torch.cat(r1[i]).clone() for i in range(4)] + [r1[4].clone()]
    r2 = M.get_matched_features(ta, tb, 500.0, 50)
    for x, y in zip(c1, [torch.cat(r2[i]) for i in range(4)] + [r2[4]]):
        assert torch.equal(x, y)


def test_tower_list_layouts_agree(M):
    """Separate (non-adjacent) tower tensors go through torch.cat, chunked views are zero-copy: same bits."""
    A, B = mo.synth_embeddings(64, 256, 1), mo.synth_embeddings(64, 256, 2)
    ta, tb = towers(A, 4), towers(B, 4)
    sep_a, sep_b = [t.clone() for t in ta], [t.clone() for t in tb]
    r1 = M.get_matched_features(ta, tb, 500.0, 20)
    r2 = M.get_matched_features(sep_a, sep_b, 500.0, 20)
    for i in range(4):
        assert torch.equal(torch.cat(r1[i]), torch.cat(r2[i]))


def test_single_batch_vs_oracle(M):
    N, D, G = 96, 1024, 3
    A, B = mo.synth_embeddings(N, D, 1, "clustered", sigma=1.0), mo.synth_embeddings(N, D, 2, "clustered", sigma=1.0)
    fa, fb = list(np.split(A, G)), list(np.split(B, G))
    ref = mo.get_matched_features_single_batch(fa, fb, 500.0, 50)
    tol = tol_f(ref[:4], mo.get_matched_features_single_batch(fa, fb, 500.0, 50, np.float32)[:4])
    got = M.get_matched_features_single_batch(towers(A, G), towers(B, G), 500.0, 50)
    for i in range(4):
        assert relerr(torch.cat(got[i]), np.concatenate(ref[i])) < tol, (i, tol)
    assert abs(float(got[4]) - ref[4]) <= TOL_ENT * abs(ref[4])


def test_random_matching(M):
    fa = [torch.full((2, 3), float(i), device="cuda") for i in range(4)]
    fb = [torch.full((2, 3), 10.0 + i, device="cuda") for i in range(4)]
    aa, bb, ab, ba, e = M.get_matched_features_random(fa, fb)
    assert [int(x[0, 0]) for x in aa] == [1, 2, 3, 0] and [int(x[0, 0]) for x in bb] == [11, 12, 13, 10]
    assert ab is fb and ba is fa and float(e) == 0.0


def test_toy_matching_cpu_mirror():
    from otgan_b200.toy_example import matching_cpu as T
    rng = np.random.RandomState(5)
    A = rng.randn(512, 16).astype(np.float32)                 # notebook-2 shape: batch 512, D=16, lambda 50, T 10
    B = (rng.randn(512, 16) * 0.5 + 1.0).astype(np.float32)
    ref = mo.toy_get_matched_features(A, B, 50.0, 10)
    got = T.get_matched_features(dev(A), dev(B), 50.0, 10)
    for i in range(4):
        assert relerr(got[i], ref[i]) < TOL_F
    assert abs(float(got[4]) - ref[4]) <= TOL_ENT * abs(ref[4])
    d = T.calc_distance(dev(A), dev(B), got)
    assert abs(float(d) - mo.toy_calc_distance(A, B, ref)) < 1e-6


@pytest.mark.parametrize("path", sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if not os.path.basename(p).startswith("ref_")))
def test_golden_vectors_on_gpu(M, path):
    g = np.load(path)
    name = os.path.basename(path)
    A, B, lam, T, G = g["A"], g["B"], float(g["lam"]), int(g["T"]), int(g["G"])
    if name.startswith("toy"):
        from otgan_b200.toy_example import matching_cpu as Tm
        got = Tm.get_matched_features(dev(A), dev(B), lam, T)
        dist = Tm.calc_distance(dev(A), dev(B), got)
        cat = lambda x: x
    elif name.startswith("single"):
        ta, tb = towers(A, G), towers(B, G)
        got = M.get_matched_features_single_batch(ta, tb, lam, T)
        dist = M.calc_distance(ta, tb, got)
        cat = torch.cat
    else:
        ta, tb = towers(A, G), towers(B, G)
        got = M.get_matched_features(ta, tb, lam, T)
        dist = M.calc_distance(ta, tb, got)
        cat = torch.cat
    tol = TOL_F
    if not name.startswith("toy"):                     # SURVEY 8d gate: max(1e-5, 1.5 x the fp32 oracle's own error on these inputs)
        fa, fb = list(np.split(A, G)), list(np.split(B, G))
        fn32 = mo.get_matched_features_single_batch if name.startswith("single") else mo.get_matched_features
        tol = tol_f([g[k] for k in ("f_aa", "f_bb", "f_ab", "f_ba")], fn32(fa, fb, lam, T, np.float32)[:4])
    for i, k in enumerate(["f_aa", "f_bb", "f_ab", "f_ba"]):
        assert relerr(cat(got[i]), g[k]) < tol, (k, tol)
    # small-h fixtures: the entropy is a mean over a handful of rows, so use an absolute floor
    assert abs(float(got[4]) - float(g["entropy"])) <= TOL_ENT * max(abs(float(g["entropy"])), 0.1)
    assert abs(float(dist) - float(g["dist"])) < TOL_DIST


def test_errors_surface_as_exceptions(M):
    from otgan_b200 import _lib
    with pytest.raises(ValueError):
        M.get_matched_features([torch.zeros(2, 4, device="cuda")] * 3, [torch.zeros(2, 4, device="cuda")] * 3, 1.0, 1)
    with pytest.raises(TypeError):
        M.get_matched_features([torch.zeros(2, 4, device="cuda", dtype=torch.float64)] * 2,
                               [torch.zeros(2, 4, device="cuda", dtype=torch.float64)] * 2, 1.0, 1)


def test_results_are_caller_owned(M):
    """A second call must not overwrite the tensors the first call returned (round-1 hazard: shared persistent buffers)."""
    rng = np.random.RandomState(3)
    X1, X2 = dev(rng.rand(16, 64).astype(np.float32)), dev(rng.rand(16, 64).astype(np.float32))
    L1 = M.cost_blocks([X1], [X1], 10.0)
    keep = L1.clone()
    L2 = M.cost_blocks([X2], [X2], 10.0)
    assert L1.data_ptr() != L2.data_ptr() and torch.equal(L1, keep)
    P1, e1, _ = M.sinkhorn(L1, 10.0, 5)
    keepP = P1.clone()
    P2, e2, _ = M.sinkhorn(L2, 10.0, 5)
    assert P1.data_ptr() != P2.data_ptr() and torch.equal(P1, keepP) and e1.data_ptr() != e2.data_ptr()


@pytest.mark.parametrize("rows,cols,T", [(1100, 1100, 3), (1300, 2500, 2)])
def test_sinkhorn_any_size(M, rows, cols, T):
    """Blocks above 1024 (the reference's default flags give h = 2500): the online-LSE streaming kernels."""
    rng = np.random.RandomState(rows)
    C = rng.rand(1, rows, cols)
    L0 = dev((-50.0 * C).astype(np.float32))
    P, ent, pc = M.sinkhorn(L0, 50.0, T)
    C32 = L0.cpu().double().numpy() / -50.0
    p, e, _ = mo.sinkhorn(C32[0], 50.0, T, np.float64)
    assert relerr(P[0], p) < TOL_P and abs(float(ent[0]) - e) <= TOL_ENT * abs(e)
    assert abs(float(pc[0]) - np.sum(p * C32[0])) < 2e-5 * rows


def test_row_restricted_feature_gradients_are_bitwise_the_full_ones(M):
    """otgan_grad_features_rows_f32 (what one data-parallel rank computes): the rows of the requested range equal the full
    computation bit for bit (same tiles, same order), for ranges inside one half and at h = 256 inside one row tile."""
    from otgan_b200 import _lib
    lib = _lib.load()
    for h, D, lo, hi in ((128, 4096, 32, 64), (128, 4096, 160, 192), (256, 2048, 128, 192), (256, 2048, 320, 384), (64, 512, 0, 128)):
        rng = np.random.RandomState(h + lo)
        P = rng.rand(6, h, h)
        P /= P.sum(axis=2, keepdims=True)
        Pd = dev(P.astype(np.float32))
        Ad, Bd = dev(rng.rand(2 * h, D).astype(np.float32)), dev(rng.rand(2 * h, D).astype(np.float32))
        ws, wsb = M._plan_ws(Ad.device, h)
        s = torch.cuda.current_stream().cuda_stream
        Ga, Gb = torch.empty(2 * h, D, device="cuda"), torch.empty(2 * h, D, device="cuda")
        _lib.check(lib.otgan_grad_features_f32(h, D, Pd.data_ptr(), Ad.data_ptr(), Bd.data_ptr(), D, Ga.data_ptr(), Gb.data_ptr(), D,
                                               ws.data_ptr(), wsb, 0, s), "grad_features")
        Ra, Rb = torch.full((2 * h, D), float("nan"), device="cuda"), torch.full((2 * h, D), float("nan"), device="cuda")
        _lib.check(lib.otgan_grad_features_rows_f32(h, D, Pd.data_ptr(), Ad.data_ptr(), Bd.data_ptr(), D, Ra.data_ptr(), Rb.data_ptr(), D,
                                                    lo, hi, ws.data_ptr(), wsb, 0, s), "grad_features_rows")
        assert torch.equal(Ra[lo:hi], Ga[lo:hi]) and torch.equal(Rb[lo:hi], Gb[lo:hi]), (h, lo, hi)


def test_cost_row_slabs_match_the_full_blocks(M):
    """The per-rank row slabs of partition "S" ([rows, h] against the full column batch) against the full h x h blocks."""
    h, D, bs = 256, 4096, 64
    A, B = mo.synth_embeddings(2 * h, D, 1, "clustered", sigma=1.0), mo.synth_embeddings(2 * h, D, 2, "clustered", sigma=1.0)
    Ad, Bd = dev(A), dev(B)
    a1, a2, b1, b2 = Ad[:h], Ad[h:], Bd[:h], Bd[h:]
    full = M.cost_blocks([a1, b2, a1, a1, a2, a2], [a2, b1, b1, b2, b1, b2], 500.0)
    for lo in (0, 128, 192):
        xa = Ad[lo:lo + bs]
        slab = M.cost_blocks([xa, xa, xa], [a2, b1, b2], 500.0)
        for k, blk in enumerate((0, 2, 3)):
            assert float((slab[k] - full[blk, lo:lo + bs]).abs().max()) / 500.0 < 1e-6
    xa, xb = Ad[h + 64:h + 128], Bd[h + 64:h + 128]
    slab = M.cost_blocks([xa, xa, xb], [b1, b2, b1], 500.0)
    for k, blk in enumerate((4, 5, 1)):
        assert float((slab[k] - full[blk, 64:128]).abs().max()) / 500.0 < 1e-6
