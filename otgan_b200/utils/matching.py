"""Drop-in host mirror of the reference's ``utils/matching.py`` on top of libotgan.so (sm_100a CUDA).

Same function names, argument meaning and return structure as the reference
(/root/reference/utils/matching.py): features are Python lists of G per-tower tensors ``[bs, D]`` (torch CUDA fp32
instead of TF graph tensors), results are four lists of G tensors ``[bs, D]`` plus a scalar (0-dim tensor) entropy.

    get_matched_features_random(features_a, features_b)                              utils/matching.py:3-9
    get_matched_features(features_a, features_b, sinkhorn_lambda, nr_sinkhorn_iter)  utils/matching.py:11-85
    get_matched_features_single_batch(...)                                           utils/matching.py:88-136
    calc_distance(features_a, features_b, matched_features)                          utils/matching.py:139-153

plus the fused entry point the re-hosted train loop uses (`matching_step`), which returns what train.py:101-128 derives
from the matched features -- the distance, the entropy and the two ``grad_ys`` lists -- without materialising the four
matched-feature tensors.

All compute runs in this library's CUDA kernels, asynchronously on torch's current stream; nothing here falls back to
torch ops or the CPU.  Inputs must be CUDA float32.
"""
import ctypes

import torch

from .. import _lib

__all__ = ["get_matched_features_random", "get_matched_features", "get_matched_features_single_batch",
           "calc_distance", "matching_step", "cost_blocks", "sinkhorn"]

_buffers = {}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _buf(key, shape, device, dtype=torch.float32):
    """Persistent INTERNAL scratch (split-K partials, pre-split plan planes), one per (device, stream): never handed to the
    caller, so a later call cannot overwrite an earlier call's result; calls on different streams do not share scratch."""
    k = (key, tuple(shape), device.index, dtype, torch.cuda.current_stream(device).cuda_stream)
    t = _buffers.get(k)
    if t is None:
        t = torch.empty(shape, device=device, dtype=dtype)
        _buffers[k] = t
    return t


def _plan_ws(device, h=128):
    """Scratch for the pre-split plan planes of the tensor-core plan-apply kernel (block side h)."""
    n = _lib.load().otgan_workspace_bytes_plan_h(int(h))
    return _buf("plan_ws", ((n + 3) // 4,), device), n


def _check_features(features_a, features_b):
    if len(features_a) != len(features_b) or len(features_a) == 0:
        raise ValueError("features_a and features_b must be non-empty lists of equal length")
    ref = features_a[0]
    for t in list(features_a) + list(features_b):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 2):
            raise TypeError("features must be 2-D CUDA float32 tensors (no CPU fallback exists)")
        if t.shape != ref.shape or t.device != ref.device:
            raise ValueError("all per-tower feature tensors must have the same shape and device")


def _gather(features):
    """tf.concat(features, 0) (utils/matching.py:16-19) -- zero-copy when the towers are consecutive row slabs of one
    buffer (e.g. torch.chunk of the critic's output), else one torch.cat."""
    f0 = features[0]
    bs, D = f0.shape
    if len(features) == 1 and f0.is_contiguous():
        return f0
    step = bs * D * f0.element_size()
    if all(t.is_contiguous() and t.data_ptr() == f0.data_ptr() + i * step for i, t in enumerate(features)):
        try:
            return f0.as_strided((len(features) * bs, D), (D, 1))
        except RuntimeError:
            pass
    return torch.cat([t.contiguous() for t in features], dim=0)


def cost_blocks(X, Y, lam, cost_kind=_lib.COST_COSINE, diag_add=None, impl=_lib.IMPL_AUTO):
    """L[k] = -lam * (cost(X[k], Y[k]) + diag_add[k] I) for lists of [rows, D] / [cols, D] tensors -> [nblk, rows, cols]."""
    lib = _lib.load()
    nblk = len(X)
    rows, D = X[0].shape
    cols = Y[0].shape[0]
    dev = X[0].device
    for t in list(X) + list(Y):
        assert t.is_cuda and t.dtype == torch.float32 and t.stride(1) == 1
    ldx, ldy = X[0].stride(0), Y[0].stride(0)
    assert all(t.stride(0) == ldx for t in X) and all(t.stride(0) == ldy for t in Y)
    ws_bytes = lib.otgan_workspace_bytes_cost(nblk, rows, cols, D, impl)
    ws = _buf("cost_ws", ((ws_bytes + 3) // 4,), dev)
    L = torch.empty((nblk, rows, cols), device=dev, dtype=torch.float32)      # caller-owned result (fresh every call)
    diag = _lib.float_array(diag_add) if diag_add is not None else None
    rc = lib.otgan_cost_blocks_f32(nblk, rows, cols, D, _lib.ptr_array([t.data_ptr() for t in X]),
                                   _lib.ptr_array([t.data_ptr() for t in Y]), ldx, ldy, cost_kind, diag,
                                   float(lam), L.data_ptr(), ws.data_ptr(), ws_bytes, impl, _stream())
    _lib.check(rc, "otgan_cost_blocks_f32")
    return L


def sinkhorn(L, lam, nr_iter, want_plan=True, impl=_lib.IMPL_AUTO, want_stats=False):
    """T Sinkhorn iterations on every block of L = -lam*C -> (P, entropy[nblk], pc[nblk]) (+ slow_steps[nblk] ints:
    how many half-steps of the scaling-form kernel fell back to the max-subtracted log-domain path)."""
    lib = _lib.load()
    nblk, rows, cols = L.shape
    assert L.is_cuda and L.dtype == torch.float32 and L.is_contiguous()
    # results are fresh tensors owned by the caller; the streaming rung (side > 512, or side > 128 with IMPL_SIMT) needs P as
    # working storage, the register-resident kernels (one CTA up to 128, one 8-CTA cluster up to 512) do not
    side = max(rows, cols)
    streaming = side > 512 or (side > 128 and impl != _lib.IMPL_AUTO)
    P = torch.empty((nblk, rows, cols), device=L.device, dtype=torch.float32) if (want_plan or streaming) else None
    ent = torch.empty((nblk,), device=L.device, dtype=torch.float32)
    pc = torch.empty((nblk,), device=L.device, dtype=torch.float32)
    slow = torch.zeros((nblk,), device=L.device, dtype=torch.int32) if want_stats else None
    rc = lib.otgan_sinkhorn_ex_f32(nblk, rows, cols, int(nr_iter), float(lam), L.data_ptr(),
                                   P.data_ptr() if P is not None else None, ent.data_ptr(), pc.data_ptr(),
                                   slow.data_ptr() if slow is not None else None, impl, _stream())
    _lib.check(rc, "otgan_sinkhorn_f32")
    if want_stats:
        return P, ent, pc, slow
    return P, ent, pc


def sharded_cost_blocks(A, B, h, lam, rows, rank, world, cost_kind=_lib.COST_COSINE, impl=_lib.IMPL_AUTO):
    """The six cost blocks with the ROWS sharded over the data-parallel ranks -- the reference's own partition
    (utils/matching.py:29-39: tower i computes `1 - features_a[i] . batch^T`, i.e. the row slab of its own rows) followed by its
    `tf.concat(dist, 0)` (:41-43) as ONE all-gather of the [3, rows, h] slabs (O(h^2) bytes).  A, B: the gathered [2h, D]
    embeddings; rows = (lo, hi): this rank's rows.  Ranks of the first half own rows of a1 (blocks a1a2, a1b1, a1b2), ranks of the
    second half rows of a2 and b2 (blocks a2b1, a2b2, b2b1): three slabs per rank, perfectly balanced.  Every rank ends up with
    the same bits in L [6, h, h]."""
    import torch.distributed as dist
    lo, hi = rows
    bs = hi - lo
    if world % 2 or (world // 2) * bs != h:
        raise ValueError("sharded cost blocks need an even number of ranks that tile both halves of the batch")
    a2, b1, b2 = A[h:], B[:h], B[h:]
    if lo < h:
        xa = A[lo:hi]
        Lloc = cost_blocks([xa, xa, xa], [a2, b1, b2], lam, cost_kind, None, impl)            # rows of blocks 0, 2, 3
    else:
        xa, xb = A[lo:hi], B[lo:hi]
        Lloc = cost_blocks([xa, xa, xb], [b1, b2, b1], lam, cost_kind, None, impl)            # rows of blocks 4, 5, 1
    buf = torch.empty((world, 3, bs, h), device=A.device, dtype=torch.float32)
    dist.all_gather_into_tensor(buf, Lloc)
    halves = buf.view(2, world // 2, 3, bs, h).permute(0, 2, 1, 3, 4).reshape(2, 3, h, h)
    L = torch.empty((6, h, h), device=A.device, dtype=torch.float32)
    for src, blocks in ((halves[0], (0, 2, 3)), (halves[1], (4, 5, 1))):      # plain copies: capturable in a CUDA graph
        for k, blk in enumerate(blocks):
            L[blk].copy_(src[k])
    return L


def _two_batch_plans(features_a, features_b, sinkhorn_lambda, nr_sinkhorn_iter, cost_kind=_lib.COST_COSINE,
                     impl=_lib.IMPL_AUTO, shard=None):
    """Cost blocks + Sinkhorn of utils/matching.py:11-61.  Returns (A, B, h, P[6,h,h], ent[6], pc[6]).
    shard = (rows, rank, world): compute only this rank's row slabs of the cost blocks and all-gather them."""
    _check_features(features_a, features_b)
    ngpu = len(features_a)
    if ngpu % 2 != 0:
        raise ValueError("get_matched_features needs an even number of towers (train.py:34)")
    A, B = _gather(features_a), _gather(features_b)
    h = A.shape[0] // 2
    a1, a2, b1, b2 = A[:h], A[h:], B[:h], B[h:]
    if shard is not None:
        L = sharded_cost_blocks(A, B, h, sinkhorn_lambda, shard[0], shard[1], shard[2], cost_kind, impl)
        P, ent, pc = sinkhorn(L, sinkhorn_lambda, nr_sinkhorn_iter, True, impl)
        return A, B, h, P, ent, pc
    # block order of utils/matching.py:41-43: a1a2, b2b1, a1b1, a1b2, a2b1, a2b2
    L = cost_blocks([a1, b2, a1, a1, a2, a2], [a2, b1, b1, b2, b1, b2], sinkhorn_lambda, cost_kind, None, impl)
    P, ent, pc = sinkhorn(L, sinkhorn_lambda, nr_sinkhorn_iter, True, impl)
    return A, B, h, P, ent, pc


def get_matched_features_random(features_a, features_b):
    """utils/matching.py:3-9 (--no_sinkhorn ablation): rotate the tower lists by one, zero entropy."""
    features_a_a = features_a[1:] + features_a[:1]
    features_b_b = features_b[1:] + features_b[:1]
    return features_a_a, features_b_b, features_b, features_a, torch.zeros((), device=features_a[0].device)


def get_matched_features(features_a, features_b, sinkhorn_lambda, nr_sinkhorn_iter, impl=_lib.IMPL_AUTO):
    """utils/matching.py:11-85."""
    lib = _lib.load()
    A, B, h, P, ent, _pc = _two_batch_plans(features_a, features_b, sinkhorn_lambda, nr_sinkhorn_iter, impl=impl)
    ngpu = len(features_a)
    N, D = A.shape
    outs = [torch.empty((N, D), device=A.device, dtype=torch.float32) for _ in range(4)]
    ws, ws_bytes = _plan_ws(A.device, h)
    rc = lib.otgan_matched_two_batch_f32(h, D, P.data_ptr(), A.data_ptr(), B.data_ptr(), A.stride(0),
                                         outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(),
                                         D, ws.data_ptr(), ws_bytes, impl, _stream())
    _lib.check(rc, "otgan_matched_two_batch_f32")
    entropy = ent.sum() / 6.0       # sum(entropy)/len(entropy), utils/matching.py:61
    f_aa, f_bb, f_ab, f_ba = (list(torch.chunk(o, ngpu, dim=0)) for o in outs)
    return f_aa, f_bb, f_ab, f_ba, entropy


def get_matched_features_single_batch(features_a, features_b, sinkhorn_lambda, nr_sinkhorn_iter, impl=_lib.IMPL_AUTO):
    """utils/matching.py:88-136: three N x N blocks, +999 on the aa/bb diagonals."""
    lib = _lib.load()
    _check_features(features_a, features_b)
    ngpu = len(features_a)
    A, B = _gather(features_a), _gather(features_b)
    N, D = A.shape
    L = cost_blocks([A, B, A], [A, B, B], sinkhorn_lambda, _lib.COST_COSINE, [999.0, 999.0, 0.0], impl)
    P, ent, _pc = sinkhorn(L, sinkhorn_lambda, nr_sinkhorn_iter, True, impl)
    outs = [torch.empty((N, D), device=A.device, dtype=torch.float32) for _ in range(4)]
    ws, ws_bytes = _plan_ws(A.device, N)
    rc = lib.otgan_matched_single_batch_f32(N, D, P.data_ptr(), A.data_ptr(), B.data_ptr(), A.stride(0),
                                            outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
                                            outs[3].data_ptr(), D, ws.data_ptr(), ws_bytes, impl, _stream())
    _lib.check(rc, "otgan_matched_single_batch_f32")
    entropy = ent[:3].sum() / 3.0
    f_aa, f_bb, f_ab, f_ba = (list(torch.chunk(o, ngpu, dim=0)) for o in outs)
    return f_aa, f_bb, f_ab, f_ba, entropy


def calc_distance(features_a, features_b, matched_features):
    """utils/matching.py:139-153: (sum b*f_bb + sum a*f_aa - 2 sum a*f_ab) / (2 * batch_size * ngpu)."""
    lib = _lib.load()
    _check_features(features_a, features_b)
    ngpu = len(features_a)
    bs, D = features_a[0].shape
    f_aa, f_bb, f_ab, _f_ba, _ = matched_features
    A, B = _gather(features_a), _gather(features_b)
    Faa, Fbb, Fab = _gather(list(f_aa)), _gather(list(f_bb)), _gather(list(f_ab))
    n = A.shape[0]
    ld = A.stride(0)
    for t in (B, Faa, Fbb, Fab):
        if t.stride(0) != ld:
            raise ValueError("calc_distance: mismatched row strides")
    ws_bytes = lib.otgan_workspace_bytes_distance(n, D)
    ws = _buf("dist_ws", ((ws_bytes + 3) // 4,), A.device)
    out = torch.empty((1,), device=A.device, dtype=torch.float32)
    rc = lib.otgan_calc_distance_f32(n, D, A.data_ptr(), B.data_ptr(), Faa.data_ptr(), Fbb.data_ptr(), Fab.data_ptr(),
                                     ld, 1.0 / (2.0 * bs * ngpu), out.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
    _lib.check(rc, "otgan_calc_distance_f32")
    return out[0]


def matching_step(features_a, features_b, sinkhorn_lambda, nr_sinkhorn_iter, impl=_lib.IMPL_AUTO, rows=None, shard=None):
    """Fused form of train.py:96-128 for the two-batch matching: returns (grad_a list, grad_b list, stats) where
    grad_a[i] = f_aa[i] - f_ab[i] (grad_ys of the fake features, train.py:111), grad_b[i] = f_bb[i] - f_ba[i]
    (grad_ys of the real features, train.py:126) and stats is a 2-element CUDA tensor [distance, entropy]
    (calc_distance via the <P,C> identity; utils/matching.py:61).  rows = (lo, hi): only these rows of grad_a / grad_b are
    needed (a data-parallel rank's own towers) -- the half-blocks / row tiles outside the range are skipped and their rows left
    undefined.  shard = (rank, world): additionally shard the cost blocks' rows over the ranks (sharded_cost_blocks)."""
    lib = _lib.load()
    A, B, h, P, ent, pc = _two_batch_plans(features_a, features_b, sinkhorn_lambda, nr_sinkhorn_iter, impl=impl,
                                           shard=(rows, shard[0], shard[1]) if (shard is not None and rows is not None) else None)
    ngpu = len(features_a)
    N, D = A.shape
    Ga = torch.empty((N, D), device=A.device, dtype=torch.float32)
    Gb = torch.empty((N, D), device=A.device, dtype=torch.float32)
    ws, ws_bytes = _plan_ws(A.device, h)
    if rows is None:
        rc = lib.otgan_grad_features_f32(h, D, P.data_ptr(), A.data_ptr(), B.data_ptr(), A.stride(0), Ga.data_ptr(),
                                         Gb.data_ptr(), D, ws.data_ptr(), ws_bytes, impl, _stream())
    else:
        rc = lib.otgan_grad_features_rows_f32(h, D, P.data_ptr(), A.data_ptr(), B.data_ptr(), A.stride(0), Ga.data_ptr(),
                                              Gb.data_ptr(), D, int(rows[0]), int(rows[1]), ws.data_ptr(), ws_bytes, impl, _stream())
    _lib.check(rc, "otgan_grad_features_f32")
    stats = torch.empty((2,), device=A.device, dtype=torch.float32)
    rc = lib.otgan_distance_from_pc_f32(pc.data_ptr(), ent.data_ptr(), N, stats.data_ptr(), _stream())
    _lib.check(rc, "otgan_distance_from_pc_f32")
    return list(torch.chunk(Ga, ngpu, dim=0)), list(torch.chunk(Gb, ngpu, dim=0)), stats
