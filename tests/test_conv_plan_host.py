"""CPU tests of the HOST logic of the tcgen05 convolution kernels (conv_tc.cu) -- no GPU, no driver.

otgan_conv_plan_describe runs the real launch functions up to the point of the kernel launch (tensor-map encoding skipped)
and returns the kernel parameters: pixel boxes, parity classes, filter-tap tables (activation view, pixel shift, weight
column / row), output strides, split factors, kernel variant.  These tests REPLAY those tables in numpy -- out += view(x)
shifted by the tap @ W[rows, cols] with zeros outside the view, exactly what the TMA boxes + tensor-core MMAs compute -- and
compare with tf.nn.conv2d(x, W, [1,s,s,1], 'SAME') (utils/nn.py:241) and its gradients computed by torch in float64.  A
wrong tap offset, parity class, 'SAME' padding, weight column or output stride is caught here, on the CPU."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

OPS = {"fprop": 0, "dgrad": 1, "wgrad": 2, "up2_fprop": 3, "up2_dgrad": 4, "up2_wgrad": 5}


@pytest.fixture(scope="module")
def lib():
    from otgan_b200 import _lib, build
    build.build()
    return _lib.load()


def same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def describe(lib, op, B, H, W, Cin, Cout, k, s, pad):
    buf = (ctypes.c_longlong * 1024)()
    n = lib.otgan_conv_plan_describe(OPS[op], B, H, W, Cin, Cout, k, k, s, pad, pad, buf, 1024)
    assert n > 0, lib.otgan_last_error()
    v = list(buf[:n])
    if v[0] in (0, 1):
        keys = ["kind", "TN", "n_cls", "n_items", "splits", "m_tiles", "n_tiles", "bw", "bh", "bn", "tiles_w", "tiles_h", "kchunks",
                "n_valid", "osW", "osH", "osN", "n_full", "tail_splits", "n_tail", "ntaps"]
        d = dict(zip(keys, v))
        o = len(keys)
        d["cls_tap_begin"], d["cls_out_off"] = v[o:o + 5], v[o + 5:o + 9]
        o += 9
    else:
        keys = ["kind", "TN", "ntaps", "co_tiles", "ci_tiles", "splits", "n_items", "bw", "bh", "bn", "tiles_w", "tiles_h", "nchunks",
                "chunks_per_split", "ldw", "split_stride"]
        d = dict(zip(keys, v))
        o = len(keys)
    d["taps"] = [dict(zip(("map", "dw", "dh", "wcol", "brow", "dmap"), v[o + 6 * t:o + 6 * t + 6])) for t in range(d["ntaps"])]
    assert o + 6 * d["ntaps"] == n
    return d


def view(x, s, m):
    """Parity view m = ph * s + pw of an NHWC array: x[:, ph::s, pw::s, :]  (make_view_map)."""
    return x[:, (m // s)::s, (m % s)::s, :]


def shifted(v, dh, dw):
    """v[n, i + dh, j + dw, :] with zeros outside the view -- the TMA box with signed coordinates and zero fill."""
    B, H, W, C = v.shape
    out = np.zeros_like(v)
    i0, i1 = max(0, -dh), min(H, H - dh)
    j0, j1 = max(0, -dw), min(W, W - dw)
    if i0 < i1 and j0 < j1:
        out[:, i0:i1, j0:j1] = v[:, i0 + dh:i1 + dh, j0 + dw:j1 + dw]
    return out


def replay_gemm(d, src, view_s, wmat, N, K, grid, out_numel):
    """out[cls_off + n osN + i osH + j osW + c] += sum over the class's taps of shifted(view)[n,i,j,:K] . wmat[brow + c, wcol : wcol + K]"""
    B, gh, gw = grid
    out = np.zeros(out_numel)
    n_idx, i_idx, j_idx = np.meshgrid(np.arange(B), np.arange(gh), np.arange(gw), indexing="ij")
    for c in range(d["n_cls"]):
        acc = np.zeros((B, gh, gw, N))
        for t in d["taps"][d["cls_tap_begin"][c]:d["cls_tap_begin"][c + 1]]:
            a = shifted(view(src, view_s, t["map"]), t["dh"], t["dw"])
            acc += a @ wmat[t["brow"]:t["brow"] + N, t["wcol"]:t["wcol"] + K].T
        base = d["cls_out_off"][c] + n_idx * d["osN"] + i_idx * d["osH"] + j_idx * d["osW"]
        for ch in range(N):
            np.add.at(out, (base + ch).ravel(), acc[..., ch].ravel())
    return out


def ref_conv(x, w, k, s):
    """torch float64: x NHWC, w [Cout, k, k, Cin]; TensorFlow 'SAME'."""
    pt, pb = same_pad(x.shape[1], k, s)
    pl, pr = same_pad(x.shape[2], k, s)
    return F.conv2d(F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb)), w.permute(0, 3, 1, 2), stride=s).permute(0, 2, 3, 1)


def presum(w, k, pad):
    """numpy restatement of up2_presum (same as tests/test_subpixel_identity.py): w [Cout, k, k, Cin] -> [4, Cout, n1*n1, Cin]."""
    idx = [[(a + kk - pad) // 2 - (a - pad) // 2 for kk in range(k)] for a in range(2)]
    n1 = idx[0][-1] + 1
    sub = np.zeros((4, w.shape[0], n1 * n1, w.shape[3]))
    for a in range(2):
        for b in range(2):
            for kh in range(k):
                for kw in range(k):
                    sub[2 * a + b, :, idx[a][kh] * n1 + idx[b][kw]] += w[:, kh, kw]
    return sub, idx, n1


def structural_checks(d, pixels_per_class, K, box=128):
    assert d["bw"] * d["bh"] * d["bn"] == box
    if d["kind"] in (0, 1):
        assert d["m_tiles"] * 128 == pixels_per_class and d["kchunks"] * 32 == K
        assert d["cls_tap_begin"][0] == 0 and d["cls_tap_begin"][d["n_cls"]] == d["ntaps"]
        tiles = d["m_tiles"] * d["n_tiles"] * d["n_cls"]
        if d["kind"] == 1:
            assert d["TN"] == 256 and d["n_items"] * 2 == tiles and d["splits"] == 1
        else:
            assert d["n_items"] == d["n_full"] + d["n_tail"] * d["tail_splits"] and d["n_full"] == (tiles - d["n_tail"]) * d["splits"]


@pytest.mark.parametrize("k,s", [(5, 1), (5, 2), (3, 1), (3, 2)])
def test_fprop_tap_tables_reproduce_same_convolution(lib, k, s):
    rng = np.random.RandomState(k * 10 + s)
    B, H, W, Cin, Cout = 8, 8, 8, 32, 128
    pad = same_pad(H, k, s)[0]
    d = describe(lib, "fprop", B, H, W, Cin, Cout, k, s, pad)
    structural_checks(d, B * (H // s) * (W // s), Cin)
    x, w = rng.randn(B, H, W, Cin), rng.randn(Cout, k, k, Cin)
    out = replay_gemm(d, x, s, w.reshape(Cout, -1), Cout, Cin, (B, H // s, W // s), B * (H // s) * (W // s) * Cout)
    ref = ref_conv(torch.from_numpy(x), torch.from_numpy(w), k, s).numpy()
    np.testing.assert_allclose(out.reshape(ref.shape), ref, rtol=0, atol=1e-10)


@pytest.mark.parametrize("k,s", [(5, 1), (5, 2), (3, 1), (3, 2)])
def test_dgrad_parity_classes_reproduce_the_input_gradient(lib, k, s):
    rng = np.random.RandomState(100 + k * 10 + s)
    B, H, W, Cin, Cout = 8, 8, 8, 128, 32
    pad = same_pad(H, k, s)[0]
    d = describe(lib, "dgrad", B, H, W, Cin, Cout, k, s, pad)
    assert d["n_cls"] == s * s
    if (k, s) == (5, 2):                               # the 5x5 filter splits into 2x2 / 2x3 / 3x2 / 3x3 taps over the 4 parities
        assert [d["cls_tap_begin"][c + 1] - d["cls_tap_begin"][c] for c in range(4)] == [4, 6, 6, 9]
    structural_checks(d, B * (H // s) * (W // s), Cout)
    x = torch.from_numpy(rng.randn(B, H, W, Cin)).requires_grad_(True)
    w = torch.from_numpy(rng.randn(Cout, k, k, Cin))
    dy = rng.randn(B, H // s, W // s, Cout)
    (ref,) = torch.autograd.grad([ref_conv(x, w, k, s)], [x], [torch.from_numpy(dy)])
    w_ihwo = w.numpy().reshape(Cout, k * k, Cin).transpose(2, 1, 0).reshape(Cin, -1)          # otgan_ohwi_to_ihwo_f32
    out = replay_gemm(d, dy, 1, w_ihwo, Cin, Cout, (B, H // s, W // s), B * H * W * Cin)
    np.testing.assert_allclose(out.reshape(B, H, W, Cin), ref.numpy(), rtol=0, atol=1e-10)


@pytest.mark.parametrize("k,s", [(5, 1), (5, 2), (3, 2)])
def test_wgrad_tap_tables_reproduce_the_filter_gradient(lib, k, s):
    rng = np.random.RandomState(200 + k * 10 + s)
    B, H, W, Cin, Cout = 8, 8, 8, 128, 128
    pad = same_pad(H, k, s)[0]
    d = describe(lib, "wgrad", B, H, W, Cin, Cout, k, s, pad)
    structural_checks(d, 0, 0, box=32)
    assert d["ldw"] == k * k * Cin and d["nchunks"] * 32 == B * (H // s) * (W // s) and d["n_items"] == d["co_tiles"] * d["ci_tiles"] * d["ntaps"] * d["splits"]
    x = torch.from_numpy(rng.randn(B, H, W, Cin))
    w = torch.from_numpy(rng.randn(Cout, k, k, Cin)).requires_grad_(True)
    dy = rng.randn(B, H // s, W // s, Cout)
    (ref,) = torch.autograd.grad([ref_conv(x, w, k, s)], [w], [torch.from_numpy(dy)])
    dw = np.zeros((Cout, d["ldw"]))
    for t in d["taps"]:
        a = shifted(view(x.numpy(), s, t["map"]), t["dh"], t["dw"])
        dw[t["brow"]:t["brow"] + Cout, t["wcol"]:t["wcol"] + Cin] += np.einsum("nijo,nijc->oc", view(dy, 1, t["dmap"]), a)
    np.testing.assert_allclose(dw.reshape(Cout, k, k, Cin), ref.numpy(), rtol=0, atol=1e-9)


@pytest.mark.parametrize("k", [5, 3])
def test_fused_upsample_tables_reproduce_resize_then_convolution(lib, k):
    """All three fused-upsample passes: 4 classes x n1^2 pre-summed taps on the low-resolution tensor == conv(resize(x, 2x))."""
    rng = np.random.RandomState(300 + k)
    B, Hl, Wl, Cin, Cout = 8, 4, 4, 128, 128
    pad = (k - 1) // 2
    xl = torch.from_numpy(rng.randn(B, Hl, Wl, Cin)).requires_grad_(True)
    w = torch.from_numpy(rng.randn(Cout, k, k, Cin)).requires_grad_(True)
    dy = rng.randn(B, 2 * Hl, 2 * Wl, Cout)
    y_ref = ref_conv(xl.repeat_interleave(2, 1).repeat_interleave(2, 2), w, k, 1)
    dx_ref, dw_ref = torch.autograd.grad([y_ref], [xl, w], [torch.from_numpy(dy)])
    sub, idx, n1 = presum(w.detach().numpy(), k, pad)
    assert n1 == lib.otgan_up2_subtaps(k, pad)
    slots = n1 * n1
    # forward
    d = describe(lib, "up2_fprop", B, Hl, Wl, Cin, Cout, k, 1, pad)
    assert d["n_cls"] == 4 and d["ntaps"] == 4 * slots
    structural_checks(d, B * Hl * Wl, Cin)
    out = replay_gemm(d, xl.detach().numpy(), 1, sub.reshape(4 * Cout, slots * Cin), Cout, Cin, (B, Hl, Wl), B * 4 * Hl * Wl * Cout)
    np.testing.assert_allclose(out.reshape(y_ref.shape), y_ref.detach().numpy(), rtol=0, atol=1e-10)
    # input gradient w.r.t. the low-resolution tensor: 4 * slots taps over the parity views of dy
    d = describe(lib, "up2_dgrad", B, Hl, Wl, Cin, Cout, k, 1, pad)
    assert d["n_cls"] == 1 and d["ntaps"] == 4 * slots
    sub_t = sub.transpose(0, 3, 2, 1).reshape(4 * Cin, slots * Cout)                            # per class: [Cin][slot][Cout]
    out = replay_gemm(d, dy, 2, sub_t, Cin, Cout, (B, Hl, Wl), B * Hl * Wl * Cin)
    np.testing.assert_allclose(out.reshape(B, Hl, Wl, Cin), dx_ref.numpy(), rtol=0, atol=1e-9)
    # filter gradient: gradients of the 4 sub-filters, then the un-sum (chain rule of the pre-sum)
    d = describe(lib, "up2_wgrad", B, Hl, Wl, Cin, Cout, k, 1, pad)
    assert d["ntaps"] == 4 * slots and d["ldw"] == slots * Cin
    dsub = np.zeros((4 * Cout, slots * Cin))
    for t in d["taps"]:
        a = shifted(view(xl.detach().numpy(), 1, t["map"]), t["dh"], t["dw"])
        dsub[t["brow"]:t["brow"] + Cout, t["wcol"]:t["wcol"] + Cin] += np.einsum("nijo,nijc->oc", view(dy, 2, t["dmap"]), a)
    dsub = dsub.reshape(4, Cout, slots, Cin)
    dw = np.zeros((Cout, k, k, Cin))
    for kh in range(k):
        for kw in range(k):
            for a in range(2):
                for b in range(2):
                    dw[:, kh, kw] += dsub[2 * a + b, :, idx[a][kh] * n1 + idx[b][kw]]
    np.testing.assert_allclose(dw, dw_ref.numpy(), rtol=0, atol=1e-9)


def test_launch_configuration_choices(lib):
    """Tile variant / split decisions at the DCGAN layer sizes (N = 256 step) and at the per-rank size of an 8-GPU run."""
    d = describe(lib, "fprop", 512, 8, 8, 1024, 1024, 5, 2, 1)            # critic conv2d_3, 512 images: 800 K-chunks per tile
    assert (d["kind"], d["TN"], d["n_items"]) == (1, 256, 128)            # 256 x 256 tiles
    d = describe(lib, "fprop", 512, 32, 32, 256, 256, 5, 2, 1)            # critic conv2d_1: 200 K-chunks per tile -> 128-row tiles
    assert (d["kind"], d["TN"], d["splits"], d["n_items"]) == (0, 256, 1, 1024)
    d = describe(lib, "fprop", 64, 8, 8, 1024, 1024, 5, 2, 1)             # 8-GPU per-rank batch: 32 tiles -> taps split 4 ways
    assert d["kind"] == 0 and d["m_tiles"] * d["n_tiles"] == 32 and d["splits"] == 4 and d["n_items"] == 128
    d = describe(lib, "dgrad", 64, 8, 8, 1024, 1024, 5, 2, 1)             # 128 tiles in 4 parity classes: unsplit
    assert d["n_cls"] == 4 and d["splits"] == 1
    d = describe(lib, "wgrad", 512, 32, 32, 256, 256, 5, 2, 1)            # 50 (co, ci, tap) items: split-K over the pixels
    assert d["n_items"] == 50 * d["splits"] and d["splits"] >= 8 and d["chunks_per_split"] * d["splits"] >= d["nchunks"]
    d = describe(lib, "wgrad", 512, 32, 32, 128, 128, 1, 1, 0)            # the 3-channel layers' 1x1 filter gradient: one tile
    assert d["n_items"] == d["splits"] and d["splits"] > 100              # ... cut along the pixels over (almost) all SMs
    buf = (ctypes.c_longlong * 4)()
    assert lib.otgan_conv_plan_describe(0, 8, 8, 8, 32, 128, 5, 5, 1, 2, 2, buf, 4) == -3            # OTGAN_ENOSPC
    assert lib.otgan_conv_plan_describe(0, 8, 8, 8, 3, 128, 5, 5, 1, 2, 2, buf, 4) == -4             # unsupported channels
    assert lib.otgan_conv_plan_describe(0, 8, 8, 8, 32, 128, 7, 7, 1, 3, 3, buf, 4) == -1            # 49 taps > tap table


GEOMS = [  # (B, H, W, k, s): rectangular images, 1x1 / 3x3 / 5x5 filters, boxes spanning several images or several rows
    (4, 16, 8, 5, 1), (4, 16, 8, 5, 2), (2, 32, 32, 3, 1), (16, 4, 4, 5, 1), (32, 4, 4, 3, 2), (1, 16, 8, 1, 1),
    (2, 8, 32, 5, 1), (16, 8, 4, 5, 2), (2, 32, 16, 3, 2), (64, 2, 2, 3, 1),
]


@pytest.mark.parametrize("geom", GEOMS)
def test_tap_tables_on_rectangular_and_small_geometries(lib, geom):
    """fprop, dgrad and wgrad tables over a spread of geometries (H != W, boxes covering several images, k in 1..5; 7x7 = 49 taps exceeds the 40-entry tap table and is rejected)."""
    B, H, W, k, s = geom
    rng = np.random.RandomState(sum(geom))
    pad_t, pad_l = same_pad(H, k, s)[0], same_pad(W, k, s)[0]
    assert pad_t == pad_l
    C1, C2 = 128, 128
    x = torch.from_numpy(rng.randn(B, H, W, C1)).requires_grad_(True)
    w = torch.from_numpy(rng.randn(C2, k, k, C1)).requires_grad_(True)
    dy = rng.randn(B, H // s, W // s, C2)
    y_ref = ref_conv(x, w, k, s)
    dx_ref, dw_ref = torch.autograd.grad([y_ref], [x, w], [torch.from_numpy(dy)])
    d = describe(lib, "fprop", B, H, W, C1, C2, k, s, pad_t)
    structural_checks(d, B * (H // s) * (W // s), C1)
    out = replay_gemm(d, x.detach().numpy(), s, w.detach().numpy().reshape(C2, -1), C2, C1, (B, H // s, W // s), y_ref.numel())
    np.testing.assert_allclose(out.reshape(y_ref.shape), y_ref.detach().numpy(), rtol=0, atol=1e-9)
    d = describe(lib, "dgrad", B, H, W, C1, C2, k, s, pad_t)
    structural_checks(d, B * (H // s) * (W // s), C2)
    w_ihwo = w.detach().numpy().reshape(C2, k * k, C1).transpose(2, 1, 0).reshape(C1, -1)
    out = replay_gemm(d, dy, 1, w_ihwo, C1, C2, (B, H // s, W // s), B * H * W * C1)
    np.testing.assert_allclose(out.reshape(B, H, W, C1), dx_ref.numpy(), rtol=0, atol=1e-9)
    d = describe(lib, "wgrad", B, H, W, C1, C2, k, s, pad_t)
    dw = np.zeros((C2, d["ldw"]))
    for t in d["taps"]:
        a = shifted(view(x.detach().numpy(), s, t["map"]), t["dh"], t["dw"])
        dw[t["brow"]:t["brow"] + C2, t["wcol"]:t["wcol"] + C1] += np.einsum("nijo,nijc->oc", view(dy, 1, t["dmap"]), a)
    np.testing.assert_allclose(dw.reshape(C2, k, k, C1), dw_ref.numpy(), rtol=0, atol=1e-8)


def test_python_support_predicates_agree_with_the_host_launch_code(lib):
    """nn.conv_tc_supported / conv_up2_supported / conv_narrow_supported decide in Python whether a layer goes to the tcgen05
    kernels; if one of them said yes for a shape the C++ launch code refuses, training would raise at run time.  Sweep
    batches, extents, channels, filters and strides: every shape a predicate accepts must be accepted by ALL the passes it
    routes to (checked with the launch code itself, in plan-capture mode)."""
    from otgan_b200.utils import nn
    buf = (ctypes.c_longlong * 1024)()

    def ok(op, B, H, W, Cin, Cout, k, s, pad):
        return lib.otgan_conv_plan_describe(op, B, H, W, Cin, Cout, k, k, s, pad, pad, buf, 1024) > 0

    n_tc = n_up2 = n_narrow = n_rejected = 0
    for B in (1, 2, 3, 4, 6, 8, 16, 24, 32, 64):
        for H, W in ((4, 4), (8, 8), (16, 16), (32, 32), (16, 8), (8, 32), (64, 64), (12, 12)):
            for k in (1, 3, 5):
                for s in (1, 2):
                    pad = same_pad(H, k, s)[0]
                    if pad != same_pad(W, k, s)[0] or pad >= k:
                        continue
                    for Cin, Cout in ((128, 128), (256, 128), (128, 256), (1024, 1024), (96, 128), (128, 144)):
                        if nn.conv_tc_supported((B, H, W, Cin), Cout, k, k, [s, s], "SAME"):
                            n_tc += 1
                            assert all(ok(op, B, H, W, Cin, Cout, k, s, pad) for op in (0, 1, 2)), (B, H, W, Cin, Cout, k, s)
                        else:
                            n_rejected += 1
                        if s == 1 and k % 2 == 1 and nn.conv_up2_supported((B, H, W, Cin), Cout, k, k, [1, 1], "SAME"):
                            n_up2 += 1
                            assert all(ok(op, B, H, W, Cin, Cout, k, 1, (k - 1) // 2) for op in (3, 4, 5)), (B, H, W, Cin, Cout, k)
                    if s == 1:
                        for Cin, Cout in ((3, 128), (128, 3)):
                            if nn.conv_narrow_supported((B, H, W, Cin), Cout, k, k, [1, 1], "SAME"):
                                n_narrow += 1
                                wide = 128
                                # the passes _ConvNarrow launches: padded-32 GEMM, 1x1 GEMM over [pixels, 128], 1x1 wgrad
                                if Cin <= 16:
                                    assert ok(0, B, H, W, 32, Cout, k, 1, pad)
                                else:
                                    assert ok(1, B, H, W, Cin, 32, k, 1, pad)
                                assert ok(0, B, H, W, wide, 128, 1, 1, 0) and ok(2, B, H, W, 128, wide, 1, 1, 0), (B, H, W, Cin, Cout, k)
    assert n_tc > 100 and n_up2 > 30 and n_narrow > 20 and n_rejected > 100


def test_workspace_queries_cover_what_the_launch_code_uses(lib):
    """otgan_workspace_bytes_conv_{gemm,wgrad,up2_wgrad} must be >= the partial buffers the chosen split factors need
    (wgrad REQUIRES the space; fprop / dgrad would silently run unsplit, i.e. slower, if the query under-estimated)."""
    from otgan_b200.utils import nn
    for B in (8, 32, 64, 256, 512):
        for (H, W, Cin, Cout, k, s) in ((32, 32, 256, 256, 5, 2), (16, 16, 512, 512, 5, 2), (8, 8, 1024, 1024, 5, 2),
                                       (8, 8, 1024, 1024, 5, 1), (32, 32, 128, 128, 1, 1), (16, 16, 128, 256, 3, 1)):
            pad = same_pad(H, k, s)[0]
            if not nn.conv_tc_supported((B, H, W, Cin), Cout, k, k, [s, s], "SAME"):
                continue
            Ho, Wo = H // s, W // s
            d = describe(lib, "fprop", B, H, W, Cin, Cout, k, s, pad)
            if d["splits"] > 1:
                assert lib.otgan_workspace_bytes_conv_gemm(B, Ho, Wo, Cout) >= d["splits"] * B * Ho * Wo * Cout * 4
            d = describe(lib, "dgrad", B, H, W, Cin, Cout, k, s, pad)
            if d["splits"] > 1:
                assert lib.otgan_workspace_bytes_conv_gemm(B, H, W, Cin) >= d["splits"] * B * H * W * Cin * 4
            d = describe(lib, "wgrad", B, H, W, Cin, Cout, k, s, pad)
            if d["splits"] > 1:
                assert lib.otgan_workspace_bytes_conv_wgrad(B, H, W, Cin, Cout, k, k, s) >= d["splits"] * d["split_stride"] * 4
                assert d["split_stride"] == Cout * k * k * Cin
        for (Hl, Wl, Cin, Cout) in ((4, 4, 1024, 1024), (8, 8, 512, 512), (16, 16, 256, 256)):
            if not nn.conv_up2_supported((B, Hl, Wl, Cin), Cout, 5, 5, [1, 1], "SAME"):
                continue
            d = describe(lib, "up2_wgrad", B, Hl, Wl, Cin, Cout, 5, 1, 2)
            assert d["split_stride"] == 4 * Cout * 9 * Cin
            if d["splits"] > 1:
                assert lib.otgan_workspace_bytes_conv_up2_wgrad(B, Hl, Wl, Cin, Cout, 5, 5, 2, 2) >= d["splits"] * d["split_stride"] * 4
            d = describe(lib, "up2_fprop", B, Hl, Wl, Cin, Cout, 5, 1, 2)
            if d["splits"] > 1:
                assert lib.otgan_workspace_bytes_conv_gemm(B, 2 * Hl, 2 * Wl, Cout) >= d["splits"] * B * 4 * Hl * Wl * Cout * 4
            d = describe(lib, "up2_dgrad", B, Hl, Wl, Cin, Cout, 5, 1, 2)
            if d["splits"] > 1:
                assert lib.otgan_workspace_bytes_conv_gemm(B, Hl, Wl, Cin) >= d["splits"] * B * Hl * Wl * Cin * 4
