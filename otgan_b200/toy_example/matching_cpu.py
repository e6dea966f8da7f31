"""Drop-in host mirror of the reference's ``toy_example/matching_cpu.py`` (tensor API, squared-Euclidean/n cost) on
libotgan.so.  Despite the module name (kept for drop-in compatibility) everything runs in CUDA kernels.

    get_matched_features(features_a, features_b, sinkhorn_lambda, nr_sinkhorn_iter)   toy_example/matching_cpu.py:4-95
    calc_distance(features_a, features_b, matched_features)                           toy_example/matching_cpu.py:155-164

features_a / features_b are single tensors [2h, D]; the halves are split with tf.split(x, 2, axis=0) semantics (:7-8).
"""
import torch

from .. import _lib
from ..utils import matching as _m

__all__ = ["get_matched_features", "calc_distance"]


def get_matched_features(features_a, features_b, sinkhorn_lambda, nr_sinkhorn_iter, impl=_lib.IMPL_AUTO):
    lib = _lib.load()
    A, B = features_a.contiguous(), features_b.contiguous()
    if not (A.is_cuda and A.dtype == torch.float32 and A.shape == B.shape and A.shape[0] % 2 == 0):
        raise TypeError("features must be CUDA float32 tensors [2h, D] of equal shape (no CPU fallback exists)")
    # two "towers" = the two halves, exactly the reference's tf.split
    _A, _B, h, P, ent, _pc = _m._two_batch_plans(list(torch.chunk(A, 2, 0)), list(torch.chunk(B, 2, 0)),
                                                 sinkhorn_lambda, nr_sinkhorn_iter, _lib.COST_EUCLID_MEAN, impl)
    N, D = A.shape
    outs = [torch.empty((N, D), device=A.device, dtype=torch.float32) for _ in range(4)]
    ws, ws_bytes = _m._plan_ws(A.device, h)
    rc = lib.otgan_matched_two_batch_f32(h, D, P.data_ptr(), A.data_ptr(), B.data_ptr(), D, outs[0].data_ptr(),
                                         outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(), D, ws.data_ptr(),
                                         ws_bytes, impl, _m._stream())
    _lib.check(rc, "otgan_matched_two_batch_f32")
    return outs[0], outs[1], outs[2], outs[3], ent.sum() / 6.0


def calc_distance(features_a, features_b, matched_features):
    """(mean(b*f_bb) + mean(a*f_aa) - 2 mean(a*f_ab)) / 2      toy_example/matching_cpu.py:155-164"""
    lib = _lib.load()
    f_aa, f_bb, f_ab, _f_ba, _ = matched_features
    A, B = features_a.contiguous(), features_b.contiguous()
    n, D = A.shape
    ws_bytes = lib.otgan_workspace_bytes_distance(n, D)
    ws = _m._buf("dist_ws", ((ws_bytes + 3) // 4,), A.device)
    out = torch.empty((1,), device=A.device, dtype=torch.float32)
    rc = lib.otgan_calc_distance_f32(n, D, A.data_ptr(), B.data_ptr(), f_aa.contiguous().data_ptr(),
                                     f_bb.contiguous().data_ptr(), f_ab.contiguous().data_ptr(), D,
                                     1.0 / (2.0 * n * D), out.data_ptr(), ws.data_ptr(), ws_bytes, _m._stream())
    _lib.check(rc, "otgan_calc_distance_f32")
    return out[0]
