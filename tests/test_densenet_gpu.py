"""GPU tests (-m gpu) of the generic convolution mode and the DenseNet dense-block kernels (csrc/conv_tc.cu generic mode,
csrc/dense_block.cu), through the C ABI.

Bit-exactness: integer-valued operands small enough to be exact in TF32 make every product and every fp32 partial sum exact,
so any tap / padding / channel-slice / slot / permutation mistake is a hard mismatch against the float64 torch reference.
Network level: the fused DenseNet critic / generator against the literal list graph of the reference (library rung, fp32).
"""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from otgan_b200 import _lib
    return _lib.load()


def st():
    return torch.cuda.current_stream().cuda_stream


def same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def ref_conv(x, w_ohwi, k, s, bias=None):
    """float64 reference: x [B,H,W,Ci], w [Co][k*k][Ci] -> [B,H/s,W/s,Co] with TensorFlow 'SAME' padding."""
    B, H, W, Ci = x.shape
    Co = w_ohwi.shape[0]
    pt, pb = same_pad(H, k, s)
    pl, pr = same_pad(W, k, s)
    xn = F.pad(x.double().permute(0, 3, 1, 2), (pl, pr, pt, pb))
    wn = w_ohwi.double().view(Co, k, k, Ci).permute(0, 3, 1, 2)
    y = F.conv2d(xn, wn, None if bias is None else bias.double(), stride=s)
    return y.permute(0, 2, 3, 1).contiguous()


def ints(shape, lo, hi, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(lo, hi + 1, shape, generator=g).float().cuda()


def crelu8(x):
    """[.., C] -> [.., 2C] in crelu8 slot order."""
    C = x.shape[-1]
    xb = x.reshape(*x.shape[:-1], C // 8, 8)
    return torch.cat([torch.relu(xb), torch.relu(-xb)], -1).reshape(*x.shape[:-1], 2 * C)


# (B, H, W, Cin, Cout, k, stride)
SHAPES = [(3, 8, 8, 12, 20, 3, 1), (2, 16, 16, 36, 144, 3, 2), (5, 4, 4, 400, 16, 3, 1), (1, 32, 32, 4, 32, 3, 1),
          (2, 8, 8, 576, 144, 3, 2), (3, 16, 16, 64, 300, 5, 1), (7, 8, 8, 32, 4, 3, 1), (2, 64, 64, 8, 8, 5, 2)]


@pytest.mark.parametrize("B,H,W,Ci,Co,k,s", SHAPES)
def test_generic_conv_three_passes_bit_exact(lib, B, H, W, Ci, Co, k, s):
    from otgan_b200 import _lib
    x = ints((B, H, W, Ci), -3, 3, 1)
    w = ints((Co, k * k * Ci), -2, 2, 2)
    b = ints((Co,), -4, 4, 3)
    pt, pl = same_pad(H, k, s)[0], same_pad(W, k, s)[0]
    Ho, Wo = H // s, W // s
    y = torch.full((B, Ho, Wo, Co), float("nan"), device="cuda")
    rc = lib.otgan_conv2d_fprop_ex_tf32(B, H, W, Ci, Ci, Co, Co, k, k, s, pt, pl, x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), 0, st())
    _lib.check(rc, "fprop_ex")
    yr = ref_conv(x, w, k, s, b)
    assert torch.equal(y.double(), yr), float((y.double() - yr).abs().max())
    # dgrad / wgrad against autograd of the float64 reference
    dy = ints((B, Ho, Wo, Co), -2, 2, 4)
    xd = x.double().requires_grad_(True)
    wd = w.double().requires_grad_(True)
    ref_conv(xd, wd, k, s).backward(dy.double())
    wt = torch.empty((Ci, k * k * Co), device="cuda")
    _lib.check(lib.otgan_ohwi_to_ihwo_f32(Co, k * k, Ci, w.data_ptr(), wt.data_ptr(), st()), "ohwi_to_ihwo")
    dx = torch.full((B, H, W, Ci), float("nan"), device="cuda")
    rc = lib.otgan_conv2d_dgrad_ex_tf32(B, H, W, Ci, Ci, Co, Co, k, k, s, pt, pl, dy.data_ptr(), wt.data_ptr(), dx.data_ptr(), st())
    _lib.check(rc, "dgrad_ex")
    assert torch.equal(dx.double(), xd.grad), float((dx.double() - xd.grad).abs().max())
    ws = torch.empty(lib.otgan_workspace_bytes_conv_wgrad_ex(B, H, W, Ci, Co, k, k, s) // 4 + 64, device="cuda")
    dw = torch.full((Co, k * k * Ci), float("nan"), device="cuda")
    rc = lib.otgan_conv2d_wgrad_ex_tf32(B, H, W, Ci, Ci, Co, Co, k, k, s, pt, pl, dy.data_ptr(), x.data_ptr(), dw.data_ptr(), ws.data_ptr(),
                                        ws.numel() * 4, st())
    _lib.check(rc, "wgrad_ex")
    assert torch.equal(dw.double(), wd.grad), float((dw.double() - wd.grad).abs().max())


def test_generic_conv_channel_slices_and_crelu8_epilogue(lib):
    """x is a channel PREFIX of a wider buffer, y a channel slice of another one, written as crelu8(conv + bias)."""
    from otgan_b200 import _lib
    B, H, W, ld, Ci, Co, ldy, off = 3, 8, 8, 96, 40, 16, 200, 64
    xbuf = ints((B, H, W, ld), -3, 3, 5)
    w = ints((Co, 9 * Ci), -2, 2, 6)
    b = ints((Co,), -3, 3, 7)
    ybuf = torch.full((B, H, W, ldy), -777.0, device="cuda")
    rc = lib.otgan_conv2d_fprop_ex_tf32(B, H, W, Ci, ld, Co, ldy, 3, 3, 1, 1, 1, xbuf.data_ptr(), w.data_ptr(), b.data_ptr(),
                                        ybuf.data_ptr() + 4 * off, 1, st())
    _lib.check(rc, "fprop_ex crelu8")
    yr = crelu8(ref_conv(xbuf[..., :Ci], w, 3, 1, b))
    assert torch.equal(ybuf[..., off:off + 2 * Co].double(), yr)
    mask = torch.ones(ldy, dtype=torch.bool, device="cuda")
    mask[off:off + 2 * Co] = False
    assert bool((ybuf[..., mask] == -777.0).all()), "the epilogue wrote outside its slot"


def test_crelu8_kernels_and_permutation(lib):
    from otgan_b200 import _lib
    P, C = 50, 24
    x = ints((P, C), -5, 5, 8)
    z = torch.empty((P, 2 * C), device="cuda")
    _lib.check(lib.otgan_crelu8_fwd_f32(P, C, x.data_ptr(), C, z.data_ptr(), 2 * C, st()), "crelu8_fwd")
    assert torch.equal(z, crelu8(x))
    dz = ints((P, 2 * C), -3, 3, 9)
    dx = torch.empty((P, C), device="cuda")
    _lib.check(lib.otgan_crelu8_bwd_f32(P, C, z.data_ptr(), 2 * C, dz.data_ptr(), 2 * C, dx.data_ptr(), C, st()), "crelu8_bwd")
    xd = x.double().requires_grad_(True)
    crelu8(xd).backward(dz.double())
    assert torch.equal(dx.double(), xd.grad)
    # permutation: reference order [x0, -x0, x1, -x1] over list elements (utils/nn.py:198-200) -> crelu8 order
    elems = [16, 8, 24]
    n = 2 * sum(elems)
    buf = (ctypes.c_int * n)()
    assert lib.otgan_crelu8_perm_host(len(elems), (ctypes.c_int * 3)(*elems), 1, buf, n) == n
    perm = torch.tensor(list(buf))
    xs = [ints((4, c), -5, 5, 20 + i).cpu() for i, c in enumerate(elems)]
    ref = torch.relu(torch.cat([t for x_ in xs for t in (x_, -x_)], 1))
    mine = torch.cat([crelu8(x_.cuda()).cpu() for x_ in xs], 1)
    assert torch.equal(mine, ref[:, perm])


def _one_hot_filters(L, c0, seed, G=16):
    """W_r [16][9][cin_r] with ONE +-1 entry per output channel: values never grow, every (tap, channel) index is exercised."""
    rng = np.random.RandomState(seed)
    ws = []
    for r in range(L):
        cin = 2 * (c0 + G * r)
        w = np.zeros((G, 9, cin), np.float32)
        for co in range(G):
            w[co, rng.randint(9), rng.randint(cin)] = rng.choice([-1.0, 1.0])
        ws.append(torch.from_numpy(w).cuda())
    return ws


@pytest.mark.parametrize("B,H,W,base,L", [(2, 8, 8, [32], 16), (3, 4, 4, [200], 16), (1, 16, 16, [16, 16], 5), (2, 32, 32, [144, 16], 3)])
def test_dense_block_forward_backward_bit_exact(lib, B, H, W, base, L):
    from otgan_b200 import _lib
    c0, G = sum(base), 16
    geom = _lib.DenseGeom()
    geom.B, geom.H, geom.W, geom.n_base, geom.L, geom.growth = B, H, W, len(base), L, G
    for i, c in enumerate(base):
        geom.base_ch[i] = c
    ctot = lib.otgan_dense_channels(ctypes.byref(geom))
    assert ctot == 2 * (c0 + G * L)
    xs = [ints((B, H, W, c), -2, 2, 30 + i) for i, c in enumerate(base)]
    wfs = _one_hot_filters(L, c0, 40)
    bs = [ints((G,), -1, 1, 50 + r) for r in range(L)]
    Z = torch.full((B, H, W, ctot), float("nan"), device="cuda")
    P, off = B * H * W, 0
    for x, c in zip(xs, base):
        _lib.check(lib.otgan_crelu8_fwd_f32(P, c, x.data_ptr(), c, Z.data_ptr() + 8 * off, ctot, st()), "crelu8_fwd")
        off += c
    wf_all = torch.zeros((G * L, 9, ctot), device="cuda")
    for r in range(L):
        wf_all[G * r:G * r + G, :, :2 * (c0 + G * r)] = wfs[r]
    bias_all = torch.cat(bs)
    S = torch.full((B, H, W, G * L), float("nan"), device="cuda")
    rc = lib.otgan_dense_block_fprop_tf32(ctypes.byref(geom), wf_all.data_ptr(), bias_all.data_ptr(), Z.data_ptr(), S.data_ptr(), st())
    _lib.check(rc, "dense_block_fprop")
    # float64 reference in the same (crelu8) channel order
    xd = [x.double().requires_grad_(True) for x in xs]
    wd = [w.double().requires_grad_(True) for w in wfs]
    bd = [b.double().requires_grad_(True) for b in bs]
    zr = torch.cat([crelu8(x) for x in xd], -1)
    for r in range(L):
        y = ref_conv(zr, wd[r].reshape(G, -1), 3, 1, bd[r])
        zr = torch.cat([zr, crelu8(y)], -1)
    assert float(zr.abs().max()) < 2048, "test inputs left the exactly representable range"
    assert torch.equal(Z.double(), zr.detach()), float((Z.double() - zr.detach()).abs().max())
    # backward
    dZ = ints((B, H, W, ctot), -2, 2, 60)
    zr.backward(dZ.double())
    WB = torch.empty(lib.otgan_dense_wb_floats(ctypes.byref(geom)), device="cuda")
    _lib.check(lib.otgan_dense_build_wb_f32(ctypes.byref(geom), wf_all.data_ptr(), WB.data_ptr(), st()), "build_wb")
    dY = torch.full((B, H, W, G * L), float("nan"), device="cuda")
    dbase = [torch.full((B, H, W, c), float("nan"), device="cuda") for c in base]
    dW = torch.full((G * L, 9, ctot), float("nan"), device="cuda")
    db = torch.full((G * L,), float("nan"), device="cuda")
    ws = torch.empty(lib.otgan_workspace_bytes_dense_bgrad(ctypes.byref(geom)) // 4 + 64, device="cuda")
    rc = lib.otgan_dense_block_bgrad_tf32(ctypes.byref(geom), Z.data_ptr(), dZ.data_ptr(), WB.data_ptr(), dY.data_ptr(),
                                          _lib.ptr_array([t.data_ptr() for t in dbase]), dW.data_ptr(), db.data_ptr(), ws.data_ptr(),
                                          ws.numel() * 4, st())
    _lib.check(rc, "dense_block_bgrad")
    for got, ref in zip(dbase, xd):
        assert float(ref.grad.abs().max()) < 2048
        assert torch.equal(got.double(), ref.grad), float((got.double() - ref.grad).abs().max())
    for r in range(L):
        cin = 2 * (c0 + G * r)
        assert torch.equal(dW[G * r:G * r + G, :, :cin].double(), wd[r].grad), ("dW", r)
        assert torch.equal(db[G * r:G * r + G].double(), bd[r].grad), ("db", r)


def _assign(template, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in template.named_parameters():
            if n.endswith("/g"):
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            elif n.endswith("/b"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))


@pytest.mark.parametrize("which", ["discriminator", "generator"])
def test_densenet_fused_path_matches_the_literal_list_graph(which):
    """Forward and parameter / input gradients of the fused DenseNet (dense-block kernels, TF32 operands) against the literal
    list graph of the reference on the strict-fp32 library rung."""
    from otgan_b200.models import densenet
    from otgan_b200.utils import nn
    tpl = getattr(densenet, which)
    tpl.reset()
    dev = torch.device("cuda")
    B = 4
    prev_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(3)
        if which == "discriminator":
            x = (torch.rand(B, 32, 32, 3, device=dev) * 2 - 1).requires_grad_(True)
            with torch.no_grad():
                tpl(torch.zeros(B, 32, 32, 3, device=dev) + 0.1, init=True)
            call = lambda: tpl(x)
        else:
            u = [torch.rand(B, 100, device=dev) * 2 - 1] + [torch.rand(B, s, s, 16, device=dev) * 2 - 1 for s in (8, 16, 32)]
            with torch.no_grad():
                tpl(B, init=True, device=dev)
            call = lambda: tpl(B, u=u)
        _assign(tpl, 7)
        outs = {}
        for backend in ("tcgen05", "cudnn"):
            nn.CONV_BACKEND = backend
            y = call()
            gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(11)).to(dev)
            grads = torch.autograd.grad([y], [tpl.flat] + ([x] if which == "discriminator" else []), grad_outputs=[gy])
            outs[backend] = (y.detach(), [g.detach() for g in grads])
        nn.CONV_BACKEND = "tcgen05"
        y1, g1 = outs["tcgen05"]
        y0, g0 = outs["cudnn"]
        assert float((y1 - y0).abs().max() / y0.abs().max()) < 3e-3
        for a, b in zip(g1, g0):
            rel = float((a - b).norm() / b.norm())
            cos = float((a * b).sum() / (a.norm() * b.norm()))
            assert rel < 0.1 and cos > 0.995, (rel, cos)        # TF32 operands through 52 layers; same gate as the DCGAN networks
    finally:
        nn.CONV_BACKEND = "tcgen05"
        torch.backends.cudnn.allow_tf32 = prev_tf32
        tpl.reset()


def test_densenet_step_launches_no_library_convolution():
    """A DenseNet critic + generator training step must run its convolutions on this library's kernels: count cuDNN calls by
    patching torch's conv2d entry (the library rung goes through F.conv2d)."""
    from otgan_b200 import train as T
    calls = []
    orig = F.conv2d

    def spy(*a, **k):
        calls.append(tuple(a[0].shape))
        return orig(*a, **k)

    targs = T.build_parser().parse_args(["--synthetic", "--nr_gpu", "2", "--batch_size", "4", "--nr_sinkhorn_iter", "10", "--model", "densenet"])
    tr = T.Trainer(targs, torch.device("cuda", 0))
    F.conv2d = spy
    try:
        for _ in range(2):
            kind, stats = tr.step(torch.rand(8, 32, 32, 3, device="cuda") * 2 - 1)
            d, e = stats.tolist()
            assert np.isfinite(d) and np.isfinite(e)
    finally:
        F.conv2d = orig
    assert calls == [], "library convolutions were called for input shapes %s" % (calls[:5],)


def test_dense_layer_on_the_generic_kernels_matches_cublas():
    """nn.dense (utils/nn.py:315-325; the generator's 100 -> 32768 layer) as a 1x1 convolution on the generic tcgen05 kernels:
    forward and the V / g / b gradients against the cuBLAS path (F.linear, TF32 off)."""
    from otgan_b200.models import dcgan
    from otgan_b200.utils import nn
    dev = torch.device("cuda")
    dcgan.generator.reset()
    torch.manual_seed(5)
    with torch.no_grad():
        dcgan.generator(8, init=True, device=dev)
    u = torch.rand(8, 100, device=dev) * 2 - 1
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    outs = {}
    try:
        for flag in (True, False):
            nn.DENSE_ON_TCGEN05 = flag
            y = dcgan.generator(8, u=u)
            gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(2)).to(dev)
            (g,) = torch.autograd.grad([y], [dcgan.generator.flat], [gy])
            n_dense = sum(p.numel() for n, p in dcgan.generator.named_parameters() if n.startswith("generator/dense_0/"))
            outs[flag] = (y.detach(), g.detach()[:n_dense])
    finally:
        nn.DENSE_ON_TCGEN05 = True
        torch.backends.cuda.matmul.allow_tf32 = prev
    (y1, g1), (y0, g0) = outs[True], outs[False]
    assert float((y1 - y0).abs().max() / y0.abs().max()) < 5e-3
    rel = float((g1 - g0).norm() / g0.norm())
    assert rel < 2e-2, rel
    dcgan.generator.reset()
