"""GPU tests (-m gpu) of the tcgen05 implicit-GEMM convolution kernels (otgan_conv2d_{fprop,dgrad,wgrad}_tf32) through the
C ABI, against a float64 restatement of tf.nn.conv2d(x, W, [1,s,s,1], 'SAME') (utils/nn.py:241) and its two gradients.

Two kinds of check:
  * exact: inputs are small integers (exactly representable in TF32, sums < 2^24) so the tensor-core result must equal
    the float64 reference BIT FOR BIT -- any tile / tap / padding / layout mistake shows up as a hard mismatch;
  * tolerance: N(0,1) inputs, error bounded by the TF32 operand rounding (2^-11 relative per operand, random-sign sum):
    max-abs-err <= 4e-3 * max-abs-value, written next to each assert.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (B, H, W, Cin, Cout, k, stride): every DCGAN layer class at reduced batch, plus 3x3 (DenseNet-style) taps
SHAPES = [
    (8, 8, 8, 128, 128, 5, 1),
    (8, 8, 8, 128, 128, 5, 2),
    (2, 32, 32, 128, 256, 5, 1),
    (2, 32, 32, 256, 256, 5, 2),
    (4, 16, 16, 256, 128, 5, 2),
    (8, 8, 8, 256, 512, 5, 2),
    (4, 16, 16, 128, 256, 3, 1),
    (4, 16, 16, 128, 128, 3, 2),
    (16, 8, 8, 1024, 1024, 5, 2),        # DCGAN critic layers at batch 16 (several N / channel tiles per launch)
    (16, 16, 16, 512, 512, 5, 2),
    (16, 8, 8, 1024, 1024, 5, 1),        # DCGAN generator layers
    (16, 16, 16, 512, 512, 5, 1),
    (80, 16, 16, 128, 256, 5, 1),        # 160 tiles: one full wave + a 12-tile tail that is cut into shares of the taps
    (96, 32, 32, 128, 128, 5, 2),        # stride 2 with a tail (fprop 192 tiles; dgrad 768 tiles in 4 parity classes)
]


def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _ref_conv(x, w_ohwi, bias, k, s):
    """float64 reference: x NHWC, w [Cout, k, k, Cin] -> y NHWC (TensorFlow 'SAME')."""
    B, H, W, C = x.shape
    pt, pb = _same_pad(H, k, s)
    pl, pr = _same_pad(W, k, s)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xn, w_ohwi.permute(0, 3, 1, 2), bias, stride=s)
    return y.permute(0, 2, 3, 1)


def _make(shape, exact, seed):
    B, H, W, Cin, Cout, k, s = shape
    g = torch.Generator(device="cuda").manual_seed(seed)
    if exact:
        x = torch.randint(-2, 3, (B, H, W, Cin), device="cuda", generator=g).float()
        w = torch.randint(-2, 3, (Cout, k, k, Cin), device="cuda", generator=g).float()
        b = torch.randint(-4, 5, (Cout,), device="cuda", generator=g).float()
        dy = torch.randint(-2, 3, (B, H // s, W // s, Cout), device="cuda", generator=g).float()
    else:
        x = torch.randn((B, H, W, Cin), device="cuda", generator=g)
        w = torch.randn((Cout, k, k, Cin), device="cuda", generator=g) * 0.05
        b = torch.randn((Cout,), device="cuda", generator=g)
        dy = torch.randn((B, H // s, W // s, Cout), device="cuda", generator=g)
    return x, w, b, dy


def _run_ours(shape, x, w, b, dy):
    from otgan_b200.utils import nn
    B, H, W, Cin, Cout, k, s = shape
    assert nn.conv_tc_supported((B, H, W, Cin), Cout, k, k, [s, s], "SAME")
    xr = x.clone().requires_grad_(True)
    wr = w.reshape(Cout, -1).clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    geom = (k, k, s, _same_pad(H, k, s)[0], _same_pad(W, k, s)[0])
    y = nn._ConvTC.apply(xr, wr, br, geom)
    dx, dw, db = torch.autograd.grad([y], [xr, wr, br], [dy])
    return y.detach(), dx, dw.view(Cout, k, k, Cin), db


def _run_ref(shape, x, w, b, dy):
    B, H, W, Cin, Cout, k, s = shape
    xd, wd, bd = (t.double().requires_grad_(True) for t in (x, w, b))
    y = _ref_conv(xd, wd, bd, k, s)
    dx, dw, db = torch.autograd.grad([y], [xd, wd, bd], [dy.double()])
    return y.detach(), dx, dw, db


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_kernels_exact_on_integer_inputs(shape):
    from otgan_b200 import _lib
    x, w, b, dy = _make(shape, True, 1)
    tail = shape[0] >= 80                       # the two shapes with a partial last wave: also exercise the tail-split option
    _lib.load().otgan_conv_set_option(1, 1 if tail else 0)
    try:
        ours = _run_ours(shape, x, w, b, dy)
    finally:
        _lib.load().otgan_conv_set_option(1, 0)
    ref = _run_ref(shape, x, w, b, dy)
    for name, o, r in zip(("fprop", "dgrad", "wgrad", "bias-grad"), ours, ref):
        diff = (o.double() - r).abs()
        assert float(diff.max()) == 0.0, "%s %s: %d / %d elements differ, max |diff| %g" % (
            name, shape, int((diff > 0).sum()), diff.numel(), float(diff.max()))


@pytest.mark.parametrize("shape", SHAPES[:6])
def test_conv_kernels_tf32_tolerance(shape):
    x, w, b, dy = _make(shape, False, 2)
    ours = _run_ours(shape, x, w, b, dy)
    ref = _run_ref(shape, x, w, b, dy)
    for name, o, r in zip(("fprop", "dgrad", "wgrad", "bias-grad"), ours, ref):
        err = float((o.double() - r).abs().max() / r.abs().max())
        assert err <= 4e-3, (name, shape, err)          # TF32 operands (10-bit mantissa), fp32 accumulation


def test_wgrad_split_k_is_deterministic():
    shape = (2, 32, 32, 256, 256, 5, 2)
    x, w, b, dy = _make(shape, False, 3)
    a = _run_ours(shape, x, w, b, dy)
    c = _run_ours(shape, x, w, b, dy)
    for o1, o2 in zip(a, c):
        assert torch.equal(o1, o2)


def test_unsupported_shapes_are_rejected_not_miscomputed():
    from otgan_b200 import _lib
    lib = _lib.load()
    x = torch.zeros(2 * 32 * 32 * 3, device="cuda")
    w = torch.zeros(128 * 75, device="cuda")
    y = torch.zeros(2 * 32 * 32 * 128, device="cuda")
    rc = lib.otgan_conv2d_fprop_tf32(2, 32, 32, 3, 128, 5, 5, 1, 2, 2, x.data_ptr(), w.data_ptr(), None, y.data_ptr(), None, 0, None)
    assert rc == -4 and b"Cin" in lib.otgan_last_error()


def test_dcgan_networks_every_conv_call_checked_against_float64():
    """Run the DCGAN critic + generator forward and backward on the tcgen05 kernels and check EVERY convolution call the
    networks make (forward output, input gradient, filter gradient, bias gradient) against a float64 convolution of the
    very tensors that call received.  (Comparing whole-network gradients against an fp32 run is not a usable gate: a
    TF32-sized perturbation flips the sign of a few near-zero CReLU pre-activations, which moves individual bias-gradient
    sums by ~10%; see test_dcgan_networks_agree_with_library_convolutions for that loose end-to-end bound.)"""
    from otgan_b200.models.dcgan import discriminator, generator
    from otgan_b200.utils import nn
    dev = torch.device("cuda", 0)
    calls = []
    orig_fwd, orig_bwd = nn._ConvTC.forward, nn._ConvTC.backward

    def fwd(ctx, x, wt, bias, geom, wt_ihwo=None):
        y = orig_fwd(ctx, x, wt, bias, geom, wt_ihwo)
        ctx._rec = {"x": x.detach().clone(), "wt": wt.detach().clone(), "b": None if bias is None else bias.detach().clone(),
                    "geom": geom, "y": y.detach().clone()}
        calls.append(ctx._rec)
        return y

    def bwd(ctx, dy):
        res = orig_bwd(ctx, dy)
        ctx._rec.update(dy=dy.detach().clone(), dx=res[0], dw=res[1], db=res[2])
        return res

    nn._ConvTC.forward, nn._ConvTC.backward = staticmethod(fwd), staticmethod(bwd)
    nn.WN_FUSION = False                               # record the stand-alone convolution nodes (the fused nodes: test_wn_fusion_*)
    try:
        discriminator.reset(); generator.reset()
        torch.manual_seed(3)
        with torch.no_grad():
            discriminator(torch.zeros(16, 32, 32, 3, device=dev), init=True, device=dev)
            generator(init=True, device=dev, batch_size=16)
        g = torch.Generator(device="cuda").manual_seed(5)
        x = torch.rand(16, 32, 32, 3, device=dev, generator=g) * 2 - 1
        u = torch.rand(16, 100, device=dev, generator=g) * 2 - 1
        gy = torch.randn(16, 32768, device=dev, generator=g)
        f = discriminator(x)
        torch.autograd.grad([f], [discriminator.flat], [gy])
        img = generator(batch_size=16, u=u)
        with nn.frozen_params():
            f2 = discriminator(img)
        torch.autograd.grad([f2], [generator.flat], [gy])
    finally:
        nn._ConvTC.forward, nn._ConvTC.backward = staticmethod(orig_fwd), staticmethod(orig_bwd)
        nn.WN_FUSION = True
    assert len(calls) == 3 + 3                         # critic c1-c3 on the real images, critic c1-c3 on the generated images
                                                       # (the generator's layers run on _ConvUp2TC / _ConvNarrow, tested separately)
    n_dx = n_dw = 0
    for rec in calls:
        kh, kw, s, pt, pl = rec["geom"]
        cout = rec["wt"].shape[0]
        cin = rec["x"].shape[3]
        xd = rec["x"].double().requires_grad_(True)
        wd = rec["wt"].double().view(cout, kh, kw, cin).requires_grad_(True)
        bd = rec["b"].double().requires_grad_(True)
        yr = _ref_conv(xd, wd, bd, kh, s)
        dxr, dwr, dbr = torch.autograd.grad([yr], [xd, wd, bd], [rec["dy"].double()])
        tag = (tuple(rec["x"].shape), cout, s)
        assert float((rec["y"].double() - yr).abs().max() / yr.abs().max()) <= 4e-3, ("fprop", tag)
        if rec["dx"] is not None:
            n_dx += 1
            assert float((rec["dx"].double() - dxr).abs().max() / dxr.abs().max()) <= 4e-3, ("dgrad", tag)
        if rec["dw"] is not None:
            n_dw += 1
            assert float((rec["dw"].double().view_as(dwr) - dwr).abs().max() / dwr.abs().max()) <= 4e-3, ("wgrad", tag)
            assert float((rec["db"].double() - dbr).abs().max() / dbr.abs().max()) <= 4e-3, ("bias-grad", tag)
    assert n_dx == 6 and n_dw == 3                     # the critic inside the generator step builds no filter gradients


def test_dcgan_networks_agree_with_library_convolutions():
    """Loose end-to-end bound: critic and generator forward + backward with this library's convolution kernels vs the cuDNN
    rung (fp32, TF32 off) on the same parameters, inputs and output gradients.  Outputs agree to TF32 accuracy; gradients
    are compared in the L2 norm / by direction (CReLU sign flips of near-zero pre-activations, see above)."""
    from otgan_b200.models.dcgan import discriminator, generator
    from otgan_b200.utils import nn
    dev = torch.device("cuda", 0)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    res = {}
    try:
        discriminator.reset(); generator.reset()
        torch.manual_seed(3)
        with torch.no_grad():
            discriminator(torch.zeros(16, 32, 32, 3, device=dev), init=True, device=dev)
            generator(init=True, device=dev, batch_size=16)
        g = torch.Generator(device="cuda").manual_seed(5)
        x = torch.rand(16, 32, 32, 3, device=dev, generator=g) * 2 - 1
        u = torch.rand(16, 100, device=dev, generator=g) * 2 - 1
        gy = torch.randn(16, 32768, device=dev, generator=g)
        for backend in ("tcgen05", "cudnn"):
            nn.CONV_BACKEND = backend
            xr = x.clone().requires_grad_(True)
            f = discriminator(xr)
            gd, gx = torch.autograd.grad([f], [discriminator.flat, xr], [gy])
            img = generator(batch_size=16, u=u)
            with nn.frozen_params():
                f2 = discriminator(img)
            (gg,) = torch.autograd.grad([f2], [generator.flat], [gy])
            res[backend] = (f.detach(), img.detach(), gd, gx, gg)
    finally:
        nn.CONV_BACKEND = "tcgen05"
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    names = ("critic features", "generator images", "critic parameter gradient", "critic input gradient", "generator parameter gradient")
    for i, (name, a, c) in enumerate(zip(names, res["tcgen05"], res["cudnn"])):
        if i < 2:
            assert float((a - c).abs().max() / c.abs().max()) <= 5e-3, name          # TF32 operand rounding, 4 layers deep
        else:
            rel_l2 = float((a - c).norm() / c.norm())
            cos = float(torch.dot(a.flatten(), c.flatten()) / (a.norm() * c.norm()))
            assert rel_l2 <= 0.1 and cos >= 0.995, (name, rel_l2, cos)


@pytest.mark.parametrize("shape", [(8, 32, 32, 3, 128, 5, 1), (8, 32, 32, 128, 3, 5, 1), (4, 16, 16, 3, 128, 3, 1), (4, 16, 16, 128, 3, 3, 1)])
@pytest.mark.parametrize("exact", [True, False])
def test_narrow_channel_convolutions(shape, exact):
    """The 3-channel layers (critic conv2d_0, generator conv2d_3) through _ConvNarrow: 1x1 GEMM + col2im shift for the
    narrow-output passes, im2col + 1x1 wgrad GEMM for the filter gradients, channel-padded GEMM kernel for the other two.
    Integer inputs: every pass exact; N(0,1): 4e-3 (TF32 operands)."""
    from otgan_b200.utils import nn
    B, H, W, Cin, Cout, k, s = shape
    assert nn.conv_narrow_supported((B, H, W, Cin), Cout, k, k, [1, 1], "SAME")
    x, w, b, dy = _make(shape, exact, 4)
    xr = x.clone().requires_grad_(True)
    wr = w.reshape(Cout, -1).clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    geom = (k, k, 1, _same_pad(H, k, 1)[0], _same_pad(W, k, 1)[0])
    y = nn._ConvNarrow.apply(xr, wr, br, geom)
    dx, dw, db = torch.autograd.grad([y], [xr, wr, br], [dy])
    ours = (y.detach(), dx, dw.view(Cout, k, k, Cin), db)
    ref = _run_ref(shape, x, w, b, dy)
    for name, o, r in zip(("fprop", "dgrad", "wgrad", "bias-grad"), ours, ref):
        err = float((o.double() - r).abs().max() / r.abs().max())
        if exact:
            assert err == 0.0, (name, shape, err)
        else:
            assert err <= 4e-3, (name, shape, err)


UP2_SHAPES = [  # (B, Hl, Wl, Cin, Cout, k): low-resolution input, output is [B, 2Hl, 2Wl, Cout]
    (8, 4, 4, 128, 128, 5),
    (4, 8, 8, 256, 256, 5),
    (2, 16, 16, 128, 256, 5),
    (16, 4, 4, 1024, 1024, 5),      # generator conv2d_0 at batch 16
    (16, 8, 8, 512, 512, 5),        # generator conv2d_1
    (4, 8, 8, 128, 128, 3),         # DenseNet-style 3x3
]


@pytest.mark.parametrize("shape", UP2_SHAPES)
@pytest.mark.parametrize("exact", [True, False])
def test_fused_upsample_convolution(shape, exact):
    """conv2d(resize_nearest_neighbor(x, 2x), W) on the fused sub-pixel kernels (4 parity classes x 3x3 pre-summed
    sub-filters on the low-resolution input) vs float64 upsample + convolution: forward, input gradient (w.r.t. the
    low-resolution tensor), filter gradient (after the un-sum), bias gradient.  Integer inputs: exact; N(0,1): 4e-3."""
    from otgan_b200.utils import nn
    B, Hl, Wl, Cin, Cout, k = shape
    assert nn.conv_up2_supported((B, Hl, Wl, Cin), Cout, k, k, [1, 1], "SAME")
    g = torch.Generator(device="cuda").manual_seed(9)
    if exact:
        x = torch.randint(-2, 3, (B, Hl, Wl, Cin), device="cuda", generator=g).float()
        w = torch.randint(-1, 2, (Cout, k, k, Cin), device="cuda", generator=g).float()
        b = torch.randint(-4, 5, (Cout,), device="cuda", generator=g).float()
        dy = torch.randint(-2, 3, (B, 2 * Hl, 2 * Wl, Cout), device="cuda", generator=g).float()
    else:
        x = torch.randn((B, Hl, Wl, Cin), device="cuda", generator=g)
        w = torch.randn((Cout, k, k, Cin), device="cuda", generator=g) * 0.05
        b = torch.randn((Cout,), device="cuda", generator=g)
        dy = torch.randn((B, 2 * Hl, 2 * Wl, Cout), device="cuda", generator=g)
    xr = x.clone().requires_grad_(True)
    wr = w.reshape(Cout, -1).clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    p = (k - 1) // 2
    y = nn._ConvUp2TC.apply(xr, wr, br, (k, k, p, p))
    dx, dw, db = torch.autograd.grad([y], [xr, wr, br], [dy])
    xd, wd, bd = (t.double().requires_grad_(True) for t in (x, w, b))
    xu = xd.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    yr = _ref_conv(xu, wd, bd, k, 1)
    dxr, dwr, dbr = torch.autograd.grad([yr], [xd, wd, bd], [dy.double()])
    for name, o, r in zip(("fprop", "dgrad", "wgrad", "bias-grad"), (y.detach(), dx, dw.view(Cout, k, k, Cin), db), (yr.detach(), dxr, dwr, dbr)):
        err = float((o.double() - r).abs().max() / r.abs().max())
        if exact:
            assert err == 0.0, (name, shape, err)
        else:
            assert err <= 4e-3, (name, shape, err)


def test_generator_uses_the_fused_upsample_path():
    """models/dcgan.py through nn.upsample2x / nn.glu(upsample=True): the three resize -> conv pairs run on _ConvUp2TC and
    the generated images equal the materialised-upsample path to TF32 accuracy."""
    from otgan_b200.models.dcgan import generator
    from otgan_b200.utils import nn
    dev = torch.device("cuda", 0)
    generator.reset()
    torch.manual_seed(3)
    with torch.no_grad():
        generator(init=True, device=dev, batch_size=16)
        u = torch.rand(16, 100, device=dev) * 2 - 1
        calls = []
        orig = nn._ConvUp2TCWN.forward                # the weight-norm-fused node (nn.WN_FUSION, default)

        def fwd(ctx, *a):
            calls.append(a[0].shape)
            return orig(ctx, *a)

        nn._ConvUp2TCWN.forward = staticmethod(fwd)
        try:
            img = generator(batch_size=16, u=u)
            nn.UPSAMPLE_FUSION = False
            ref = generator(batch_size=16, u=u)
        finally:
            nn.UPSAMPLE_FUSION = True
            nn._ConvUp2TCWN.forward = staticmethod(orig)
    assert [tuple(c) for c in calls] == [(16, 4, 4, 1024), (16, 8, 8, 512), (16, 16, 16, 256)]
    assert float((img - ref).abs().max() / ref.abs().max()) <= 5e-3
    generator.reset()


@pytest.mark.parametrize("cout", [3, 16])
def test_narrow_output_variant_of_the_gemm_kernel(cout):
    """otgan_conv2d_fprop_tf32 with Cout <= 16 (the TN = 16 instantiation: weight box = the real rows, masked store) and
    otgan_conv2d_dgrad_tf32 with Cin <= 16, called directly through the C ABI; integer inputs, exact."""
    from otgan_b200 import _lib
    lib = _lib.load()
    B, H, W, Cin, k, s = 4, 16, 16, 128, 5, 1
    shape = (B, H, W, Cin, cout, k, s)
    x, w, b, dy = _make(shape, True, 6)
    ref = _run_ref(shape, x, w, b, dy)
    st = torch.cuda.current_stream().cuda_stream
    y = torch.full((B, H, W, cout), float("nan"), device="cuda")
    rc = lib.otgan_conv2d_fprop_tf32(B, H, W, Cin, cout, k, k, s, 2, 2, x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), None, 0, st)
    _lib.check(rc, "fprop")
    assert float((y.double() - ref[0]).abs().max()) == 0.0
    # dgrad with a narrow INPUT side: the roles of Cin / Cout swap (dy has 128 channels, dx has `cout` channels)
    shape2 = (B, H, W, cout, 128, k, s)
    x2, w2, b2, dy2 = _make(shape2, True, 7)
    ref2 = _run_ref(shape2, x2, w2, b2, dy2)
    wt = w2.view(128, k * k, cout).permute(2, 1, 0).reshape(cout, -1).contiguous()          # IHWO [Cin, k*k*Cout]
    dx = torch.full((B, H, W, cout), float("nan"), device="cuda")
    rc = lib.otgan_conv2d_dgrad_tf32(B, H, W, cout, 128, k, k, s, 2, 2, dy2.data_ptr(), wt.data_ptr(), dx.data_ptr(), None, 0, st)
    _lib.check(rc, "dgrad")
    torch.cuda.synchronize()
    assert float((dx.double() - ref2[1]).abs().max()) == 0.0


def test_256_row_tile_variant_matches_the_128_row_kernel():
    """conv_gemm2_tc_kernel (two 128-pixel sub-tiles per CTA sharing one weight tile; chosen for launches with >= 111 of the
    bigger tiles and >= 100 K-chunks per tile) against the 128 x 256-tile kernel (itself checked against float64 above) on
    integer inputs, where both are exact: outputs must be identical.  Covers fprop, stride-1 and stride-2 dgrad (parity
    classes) and the fused-upsample fprop / dgrad."""
    from otgan_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(21)

    def ints(*shape, lo=-2, hi=3):
        return torch.randint(lo, hi, shape, device="cuda", generator=g).float()

    def both(fn, out):
        res = []
        assert lib.otgan_conv_set_option(2, 100) == 0        # admit every launch of this test to the 256-row variant
        for v in (1, 0):
            assert lib.otgan_conv_set_option(0, v) == 0
            out.fill_(float("nan"))
            _lib.check(fn(), "launch")
            torch.cuda.synchronize()
            res.append(out.clone())
        lib.otgan_conv_set_option(0, 1)
        assert not torch.isnan(res[0]).any() and torch.equal(res[0], res[1])

    def ws_for(*dims):
        n = lib.otgan_workspace_bytes_conv_gemm(*dims)
        t = torch.empty(n // 4 + 64, device="cuda")
        return t, t.data_ptr(), t.numel() * 4

    try:
        # fprop, stride 1: 512 tiles of 128 x 256, 100 K-chunks per tile
        B, H, W, Cin, Cout, k = 64, 32, 32, 128, 256, 5
        x, w, b = ints(B, H, W, Cin), ints(Cout, k * k * Cin), ints(Cout, lo=-4, hi=5)
        y = torch.empty(B, H, W, Cout, device="cuda")
        t, wp, wb = ws_for(B, H, W, Cout)
        both(lambda: lib.otgan_conv2d_fprop_tf32(B, H, W, Cin, Cout, k, k, 1, 2, 2, x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), wp, wb, st), y)
        # dgrad, stride 1 (N = Cin = 256)
        Cin, Cout = 256, 128
        dy, wt = ints(B, H, W, Cout), ints(Cin, k * k * Cout)
        dx = torch.empty(B, H, W, Cin, device="cuda")
        t, wp, wb = ws_for(B, H, W, Cin)
        both(lambda: lib.otgan_conv2d_dgrad_tf32(B, H, W, Cin, Cout, k, k, 1, 2, 2, dy.data_ptr(), wt.data_ptr(), dx.data_ptr(), wp, wb, st), dx)
        # dgrad, stride 2: four parity classes with 4 / 6 / 6 / 9 taps, 32 K-chunks per tap
        B, H, W, Cin, Cout = 128, 8, 8, 1024, 1024
        dy, wt = ints(B, H // 2, W // 2, Cout), ints(Cin, k * k * Cout, lo=-1, hi=2)
        dx = torch.empty(B, H, W, Cin, device="cuda")
        t, wp, wb = ws_for(B, H, W, Cin)
        both(lambda: lib.otgan_conv2d_dgrad_tf32(B, H, W, Cin, Cout, k, k, 2, 1, 1, dy.data_ptr(), wt.data_ptr(), dx.data_ptr(), wp, wb, st), dx)
        # fused upsample: fprop (4 classes x 9 taps, strided output rows) and dgrad (36 taps over the parity views of dy)
        B, Hl, Wl, Cin, Cout = 256, 8, 8, 512, 256
        x, w_sub, b = ints(B, Hl, Wl, Cin), ints(4, Cout, 9 * Cin, lo=-1, hi=2), ints(Cout, lo=-4, hi=5)
        y = torch.empty(B, 2 * Hl, 2 * Wl, Cout, device="cuda")
        t, wp, wb = ws_for(B, 2 * Hl, 2 * Wl, Cout)
        both(lambda: lib.otgan_conv2d_up2_fprop_tf32(B, Hl, Wl, Cin, Cout, k, k, 2, 2, x.data_ptr(), w_sub.data_ptr(), b.data_ptr(), y.data_ptr(), wp, wb, st), y)
        dy, w_sub_t = ints(B, 2 * Hl, 2 * Wl, Cout), ints(4, Cin, 9 * Cout, lo=-1, hi=2)
        dx = torch.empty(B, Hl, Wl, Cin, device="cuda")
        t, wp, wb = ws_for(B, Hl, Wl, Cin)
        both(lambda: lib.otgan_conv2d_up2_dgrad_tf32(B, Hl, Wl, Cin, Cout, k, k, 2, 2, dy.data_ptr(), w_sub_t.data_ptr(), dx.data_ptr(), wp, wb, st), dx)
    finally:
        lib.otgan_conv_set_option(0, 1)
        lib.otgan_conv_set_option(2, 500)


@pytest.mark.parametrize("which", ["discriminator", "generator"])
def test_wn_fusion_matches_the_unfused_nodes(which):
    """The weight-norm-fused convolution nodes (_ConvTCWN / _ConvUp2TCWN: W and the IHWO dgrad operand from one pass over V,
    filter gradients written in V's HWIO layout, streaming weight-norm backward) against the stand-alone nodes
    (_WeightNorm -> _ConvTC / _ConvUp2TC).  Same tcgen05 kernels for the three convolution passes, so only the fp32 summation
    order of the weight-norm reductions differs: 2e-5 of the largest gradient entry."""
    from otgan_b200.models import dcgan
    from otgan_b200.utils import nn
    dev = torch.device("cuda", 0)
    tpl = getattr(dcgan, which)
    tpl.reset()
    torch.manual_seed(4)
    B = 16
    with torch.no_grad():
        if which == "discriminator":
            tpl(torch.zeros(B, 32, 32, 3, device=dev) + 0.1, init=True)
        else:
            tpl(B, init=True, device=dev)
    with torch.no_grad():
        for n, p in tpl.named_parameters():
            if n.endswith("/g"):
                p.mul_(0.5 + torch.rand_like(p))
    x = (torch.rand(B, 32, 32, 3, device=dev) * 2 - 1).requires_grad_(True)
    u = torch.rand(B, 100, device=dev) * 2 - 1
    outs = {}
    try:
        for fused in (True, False):
            nn.WN_FUSION = fused
            y = tpl(x) if which == "discriminator" else tpl(B, u=u)
            gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(9)).to(dev)
            grads = torch.autograd.grad([y], [tpl.flat] + ([x] if which == "discriminator" else []), grad_outputs=[gy])
            outs[fused] = (y.detach(), [g.detach() for g in grads])
    finally:
        nn.WN_FUSION = True
    (y1, g1), (y0, g0) = outs[True], outs[False]
    assert float((y1 - y0).abs().max() / y0.abs().max()) < 1e-5
    for a, b in zip(g1, g0):
        assert float((a - b).abs().max() / b.abs().max()) < 2e-5
    tpl.reset()


@pytest.mark.parametrize("shape", [(8, 8, 8, 128, 128, 5, 1), (2, 32, 32, 256, 256, 5, 2), (8, 8, 8, 256, 512, 5, 2), (4, 16, 16, 128, 256, 3, 1),
                                   (80, 16, 16, 128, 256, 5, 1), (32, 32, 32, 32, 128, 5, 1)])
def test_fprop_with_fused_crelu_output_exact(shape):
    """otgan_conv2d_fprop_crelu_tf32: z = [relu(conv + b) | relu(-(conv + b))] written by the convolution's epilogue, bit for bit
    against float64 on integer inputs (incl. a launch with a tail wave and the 32-channel padded form of the critic's first
    layer); and otgan_crelu_bwd_from_activated_f32 against the CReLU backward taken from the pre-activation."""
    from otgan_b200 import _lib
    lib = _lib.load()
    B, H, W, Cin, Cout, k, s = shape
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randint(-2, 3, (B, H, W, Cin), device="cuda", generator=g).float()
    w = torch.randint(-2, 3, (Cout, k, k, Cin), device="cuda", generator=g).float()
    b = torch.randint(-4, 5, (Cout,), device="cuda", generator=g).float()
    pt, pl = _same_pad(H, k, s)[0], _same_pad(W, k, s)[0]
    z = torch.full((B, H // s, W // s, 2 * Cout), float("nan"), device="cuda")
    rc = lib.otgan_conv2d_fprop_crelu_tf32(B, H, W, Cin, Cout, k, k, s, pt, pl, x.data_ptr(), w.reshape(Cout, -1).contiguous().data_ptr(),
                                           b.data_ptr(), z.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.otgan_last_error()
    y = _ref_conv(x.double(), w.double(), b.double(), k, s)
    zr = torch.relu(torch.cat([y, -y], 3))
    assert torch.equal(z.double(), zr)
    dz = torch.randint(-3, 4, z.shape, device="cuda", generator=g).float()
    dy = torch.empty((B, H // s, W // s, Cout), device="cuda")
    rc = lib.otgan_crelu_bwd_from_activated_f32(B * (H // s) * (W // s), Cout, z.data_ptr(), dz.data_ptr(), dy.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.otgan_last_error()
    dyr = torch.where(y > 0, dz[..., :Cout].double(), torch.zeros_like(y)) - torch.where(y < 0, dz[..., Cout:].double(), torch.zeros_like(y))
    assert torch.equal(dy.double(), dyr)


def test_critic_with_fused_crelu_matches_the_separate_pass():
    """models/dcgan.py critic with nn.conv2d(crelu_out=True) taking the fused epilogue on every layer against the same network with
    the separate CReLU pass.  Batch 128: every forward launch has enough tiles that neither form is split over the filter taps,
    so both run the same kernels on the same operands in the same summation order -- features and gradients must agree to the
    last bit (a 16-image batch does not: the unfused launches are then split, and last-bit differences of an activation flip
    TF32-truncated operands of the next layer by 2^-10: 1.8e-4 on the features, ~1 % on the gradients)."""
    from otgan_b200.models import dcgan
    from otgan_b200.utils import nn
    from otgan_b200 import _lib
    dev = torch.device("cuda", 0)
    tpl = dcgan.discriminator
    tpl.reset()
    torch.manual_seed(3)
    B = 128
    with torch.no_grad():
        tpl(torch.zeros(16, 32, 32, 3, device=dev) + 0.1, init=True)
        for n, p in tpl.named_parameters():
            if n.endswith("/g"):
                p.mul_(0.5 + torch.rand_like(p))
            if n.endswith("/b"):
                p.add_(0.1 * torch.randn_like(p))
    x = (torch.rand(B, 32, 32, 3, device=dev) * 2 - 1).requires_grad_(True)
    outs, launches = {}, {}
    prev = (nn.CRELU_FUSION, nn.CRELU_FUSION_MIN_TILES)
    try:
        for fused in (True, False):
            nn.CRELU_FUSION, nn.CRELU_FUSION_MIN_TILES = fused, 1
            _lib.reset_launch_count()
            y = tpl(x)
            gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(9)).to(dev)
            grads = torch.autograd.grad([y], [tpl.flat, x], grad_outputs=[gy])
            torch.cuda.synchronize()
            launches[fused] = _lib.launch_count()
            outs[fused] = (y.detach(), [g.detach() for g in grads])
    finally:
        nn.CRELU_FUSION, nn.CRELU_FUSION_MIN_TILES = prev
    (y1, g1), (y0, g0) = outs[True], outs[False]
    # three CReLU passes folded into the producing convolutions (and their launches are never split over the filter taps: no split-reduce)
    assert launches[True] <= launches[False] - 3, launches
    assert torch.equal(y1, y0)
    for a, b in zip(g1, g0):
        assert torch.equal(a, b)
    tpl.reset()
