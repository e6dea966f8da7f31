"""CPU tests of the C-ABI boundary: libotgan.so loads, exports every symbol include/otgan.h declares, the ctypes
signatures cover them all, and argument validation returns OTGAN_EINVAL before anything touches the GPU."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "otgan.h")


def _declared():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"OTGAN_API[^;(]*?\b(otgan_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from otgan_b200 import _lib, build
    build.build()
    return _lib.load()


def test_header_declares_symbols():
    names = _declared()
    assert "otgan_cost_blocks_f32" in names and "otgan_sinkhorn_f32" in names and len(names) >= 14


def test_library_exports_every_declared_symbol(lib):
    from otgan_b200 import _lib
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.SO_PATH]).decode()
    exported = set(re.findall(r" T (otgan_\w+)", out))
    missing = [n for n in _declared() if n not in exported]
    assert not missing, "declared in include/otgan.h but not exported: %s" % missing
    for n in _declared():
        assert getattr(lib, n) is not None


def test_ctypes_signatures_cover_header(lib):
    from otgan_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    assert lib.otgan_abi_version() == 1


def test_plan_struct_layout_matches_header():
    from otgan_b200 import _lib
    # int n_out + 8 ints + 3 x (8x3 ints) + 8x3 floats
    assert ctypes.sizeof(_lib.Plan) == 4 * (1 + 8 + 3 * 24 + 24)


def test_library_is_sm100a_only():
    from otgan_b200 import _lib
    out = subprocess.check_output(["cuobjdump", "-lelf", _lib.SO_PATH]).decode()
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_invalid_arguments_are_rejected_before_launch(lib):
    from otgan_b200 import _lib
    null = ctypes.c_void_p(0)
    one = _lib.ptr_array([16])
    assert lib.otgan_cost_blocks_f32(0, 4, 4, 4, one, one, 4, 4, 0, None, 1.0, null, null, 0, 0, null) == -1
    assert b"nblk" in lib.otgan_last_error()
    assert lib.otgan_cost_blocks_f32(1, 4, 4, 8, one, one, 4, 4, 0, None, 1.0, 16, 16, 0, 0, null) == -1   # ld < D
    assert lib.otgan_cost_blocks_f32(1, 4, 4, 4, one, one, 4, 4, 7, None, 1.0, 16, 16, 0, 0, null) == -1   # cost kind
    assert lib.otgan_sinkhorn_f32(1, 0, 4, 1, 1.0, 16, null, null, null, 0, null) == -1
    assert lib.otgan_sinkhorn_f32(1, 4, 4, -1, 1.0, 16, null, null, null, 0, null) == -1
    assert lib.otgan_sinkhorn_f32(1, 4, 4, 1, 1.0, null, null, null, null, 0, null) == -1
    assert lib.otgan_sinkhorn_f32(1, 4, 4, 1, 1.0, 16, 36, null, null, 0, null) == -1                      # P is 4 bytes off L0's 16-byte phase
    assert b"16 bytes" in lib.otgan_last_error()
    assert lib.otgan_sinkhorn_f32(1, 4, 4, 1, 1.0, 18, null, null, null, 0, null) == -1                    # L0 not float-aligned
    assert lib.otgan_grad_features_f32(4, 4, null, null, null, 4, null, null, 4, null, 0, 0, null) == -1
    assert lib.otgan_calc_distance_f32(4, 8, 16, 16, 16, 16, 16, 4, 1.0, 16, 16, 1 << 20, null) == -1     # ld < D
    assert lib.otgan_distance_from_pc_f32(null, null, 4, null, null) == -1
    assert lib.otgan_workspace_bytes_cost(6, 128, 128, 32768, 0) > 0
    assert lib.otgan_workspace_bytes_cost(0, 128, 128, 32768, 0) == 0
    # convolution entry points: null / misaligned pointers, bad stride, bad geometry are refused on the host
    assert lib.otgan_conv2d_fprop_tf32(8, 8, 8, 128, 128, 5, 5, 1, 2, 2, null, null, null, null, null, 0, null) == -1
    assert b"null" in lib.otgan_last_error()
    assert lib.otgan_conv2d_fprop_tf32(8, 8, 8, 128, 128, 5, 5, 3, 2, 2, 16, 16, null, 16, null, 0, null) == -1      # stride 3
    assert lib.otgan_conv2d_fprop_tf32(8, 8, 8, 128, 128, 5, 5, 1, 2, 2, 20, 16, null, 16, null, 0, null) == -1      # x misaligned
    assert lib.otgan_conv2d_dgrad_tf32(8, 8, 8, 128, 128, 5, 5, 1, 2, 2, null, 16, 16, null, 0, null) == -1
    assert lib.otgan_conv2d_wgrad_tf32(8, 8, 8, 128, 128, 5, 5, 1, 2, 2, 16, null, 16, null, 0, null) == -1
    assert lib.otgan_conv2d_wgrad_tf32(8, 8, 8, 128, 128, 5, 5, 1, 7, 2, 16, 16, 16, null, 0, null) == -1            # pad >= k
    assert lib.otgan_conv2d_up2_fprop_tf32(8, 4, 4, 128, 128, 5, 5, 2, 2, null, 16, null, 16, null, 0, null) == -1
    assert lib.otgan_im2col_narrow_f32(8, 32, 32, 3, 5, 5, 2, 2, 0, 16, 16, 64, null) == -1                          # ldc < 75
    assert lib.otgan_col2im_narrow_f32(8, 32, 32, 17, 5, 5, 2, 2, 0, 16, 512, null, 16, null) == -1                  # C > 16
    assert lib.otgan_colsum_f32(128, 6, 16, 16, 16, 1 << 20, null) == -1                                             # C % 4
    assert lib.otgan_conv_set_option(9, 1) == -1 and lib.otgan_up2_subtaps(5, 2) == 3 and lib.otgan_up2_subtaps(7, 3) == 0
    assert lib.otgan_workspace_bytes_conv_wgrad(512, 32, 32, 256, 256, 5, 5, 2) > 256 and lib.otgan_workspace_bytes_conv_gemm(0, 1, 1, 1) == 0


def test_product_path_has_no_cpu_fallback():
    """CPU tensors must be rejected, not silently computed elsewhere; and the package must not import the oracle."""
    import torch
    from otgan_b200.utils import matching
    fa = [torch.zeros(2, 4), torch.zeros(2, 4)]
    with pytest.raises(TypeError):
        matching.get_matched_features(fa, fa, 1.0, 1)
    pkg = os.path.join(ROOT, "otgan_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, fn


def test_crelu8_permutation_and_dense_geometry_host_side():
    """Host-only entry points of the DenseNet path (no GPU needed): the crelu8 <-> reference channel permutation
    (utils/nn.py:198-200 orders a CReLU'd LIST as [x0, -x0, x1, -x1, ...]) and the dense-block geometry queries."""
    import ctypes
    import numpy as np
    from otgan_b200 import _lib
    lib = _lib.load()
    elems, taps = [32, 16, 16, 8], 3
    c2 = 2 * sum(elems)
    buf = (ctypes.c_int * (taps * c2))()
    assert lib.otgan_crelu8_perm_host(len(elems), (ctypes.c_int * len(elems))(*elems), taps, buf, taps * c2) == taps * c2
    perm = np.array(list(buf))
    assert sorted(perm.tolist()) == list(range(taps * c2))                       # a permutation, tap by tap
    xs = [np.random.RandomState(i).randn(c) for i, c in enumerate(elems)]
    ref = np.maximum(np.concatenate([t for x in xs for t in (x, -x)]), 0)        # the reference's order
    mine = np.concatenate([np.concatenate([np.maximum(x.reshape(-1, 8), 0), np.maximum(-x.reshape(-1, 8), 0)], 1).reshape(-1) for x in xs])
    for t in range(taps):
        assert np.array_equal(mine, ref[perm[t * c2:(t + 1) * c2] - t * c2])
    assert lib.otgan_crelu8_perm_host(1, (ctypes.c_int * 1)(12), 1, buf, 24) < 0          # 12 channels: not a multiple of 8
    assert lib.otgan_crelu8_perm_host(len(elems), (ctypes.c_int * len(elems))(*elems), taps, buf, 10) < 0   # buffer too small
    g = _lib.DenseGeom()
    g.B, g.H, g.W, g.n_base, g.L, g.growth = 4, 8, 8, 2, 16, 16
    g.base_ch[0], g.base_ch[1] = 144, 16
    assert lib.otgan_dense_channels(ctypes.byref(g)) == 2 * (160 + 256)
    assert lib.otgan_dense_wb_floats(ctypes.byref(g)) == 2 * (160 + 256) * 9 * 256
    assert lib.otgan_workspace_bytes_dense_bgrad(ctypes.byref(g)) > 0
    g.growth = 12
    assert lib.otgan_dense_channels(ctypes.byref(g)) < 0
