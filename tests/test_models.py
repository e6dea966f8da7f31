"""CPU tests of the re-hosted layer library and models (host logic; torch CPU tensors) against the numpy oracle
(oracle/nn_oracle.py): TF 'SAME' padding, CReLU list interleave, weight norm, NN resize, variable naming/sharing."""
import numpy as np
import pytest
import torch

from oracle import nn_oracle as no
from otgan_b200.models import dcgan, densenet
from otgan_b200.utils import nn


def table_of(template):
    return {n: p.detach().double().numpy() for n, p in template.named_parameters()}


def perturb(template, seed):
    """Move g and b away from their 1/0 initial values so that the test exercises them."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in template.named_parameters():
            if n.endswith("/g"):
                p.mul_(1.0 + 0.3 * torch.rand(p.shape, generator=gen))
            if n.endswith("/b"):
                p.add_(0.1 * torch.randn(p.shape, generator=gen))


@pytest.mark.parametrize("n,k,s,expect", [(32, 5, 1, (2, 2)), (32, 5, 2, (1, 2)), (32, 3, 1, (1, 1)), (32, 3, 2, (0, 1)),
                                          (7, 3, 2, (1, 1)), (8, 5, 2, (1, 2))])
def test_same_padding_matches_tensorflow(n, k, s, expect):
    assert nn.same_padding(n, k, s) == expect and no.same_pad(n, k, s)[1:] == expect


@pytest.mark.parametrize("stride,k,cin,cout,size", [(1, 5, 3, 8, 8), (2, 5, 6, 4, 8), (2, 3, 5, 7, 8), (1, 3, 4, 4, 5), (2, 5, 2, 3, 7)])
def test_conv2d_same_vs_oracle(stride, k, cin, cout, size):
    rng = np.random.RandomState(k * 10 + stride)
    x = rng.randn(2, size, size, cin)
    W = rng.randn(k, k, cin, cout)
    y = nn._conv2d_nhwc(torch.from_numpy(x), torch.from_numpy(W), [stride, stride], "SAME").numpy()
    np.testing.assert_allclose(y, no.conv2d_same(x, W, stride), atol=1e-12)


def test_crelu_list_interleave_and_resize():
    a, b = torch.randn(1, 2, 2, 3, dtype=torch.float64), torch.randn(1, 2, 2, 2, dtype=torch.float64)
    got = nn.apply_pre_activation([a, b], "crelu", 3).numpy()
    np.testing.assert_allclose(got, no.pre_activation([a.numpy(), b.numpy()], "crelu"), atol=0)
    assert got.shape[-1] == 10 and np.allclose(got[..., 3:6], np.maximum(-a.numpy(), 0))
    x = torch.arange(2 * 3 * 3 * 1, dtype=torch.float64).reshape(2, 3, 3, 1)
    np.testing.assert_allclose(nn.resize_nearest_neighbor(x, [6, 6]).numpy(), no.resize_nn(x.numpy(), 6, 6))
    np.testing.assert_allclose(nn.resize_nearest_neighbor(x, [5, 4]).numpy(), no.resize_nn(x.numpy(), 5, 4))


def test_dcgan_matches_oracle_and_reference_shapes():
    dcgan.discriminator.reset(); dcgan.generator.reset()
    torch.manual_seed(0)
    x = torch.rand(2, 32, 32, 3) * 2 - 1
    f0 = dcgan.discriminator(x, init=True, device="cpu")
    g0 = dcgan.generator(2, init=True, device="cpu")
    assert f0.shape == (2, 32768) and g0.shape == (2, 32, 32, 3)
    assert dcgan.discriminator.store.num_params() == 34419840 and dcgan.generator.store.num_params() == 37761926
    names = [n for n, _ in dcgan.discriminator.named_parameters()]
    assert names[:3] == ["discriminator/conv2d_0/V", "discriminator/conv2d_0/g", "discriminator/conv2d_0/b"]
    assert [n for n, _ in dcgan.generator.named_parameters()][0] == "generator/dense_0/V"
    assert tuple(dcgan.discriminator.store.get("discriminator/conv2d_1/V").shape) == (5, 5, 256, 256)            # HWIO, CReLU doubles Cin
    perturb(dcgan.discriminator, 1); perturb(dcgan.generator, 2)
    f = dcgan.discriminator(x).detach().double().numpy()
    ref = no.dcgan_discriminator(x.double().numpy(), table_of(dcgan.discriminator))
    assert np.abs(f - ref).max() / np.abs(ref).max() < 2e-5
    np.testing.assert_allclose(np.linalg.norm(f, axis=1), 1.0, atol=1e-5)
    u = torch.rand(2, 100) * 2 - 1
    g = dcgan.generator(2, u=u).detach().double().numpy()
    gref = no.dcgan_generator(u.double().numpy(), table_of(dcgan.generator))
    assert np.abs(g - gref).max() < 2e-5
    # variables are shared across calls (tf.make_template) and the EMA shadow substitutes them (utils/nn.py:89-93)
    ema = nn.ExponentialMovingAverage(0.999).attach(dcgan.generator)
    ema.shadow.mul_(0.5)
    g_ema = dcgan.generator(2, u=u, ema=ema).detach()
    assert not torch.allclose(g_ema, torch.from_numpy(g).float())
    dcgan.discriminator.reset(); dcgan.generator.reset()


def test_densenet_matches_oracle_small():
    densenet.discriminator.reset(); densenet.generator.reset()
    torch.manual_seed(1)
    x = torch.rand(1, 32, 32, 3) * 2 - 1
    kw = dict(layers_per_block=3, filters_per_layer=4)
    f0 = densenet.discriminator(x, init=True, device="cpu", **kw)
    u = [torch.rand(1, 100) * 2 - 1, torch.rand(1, 8, 8, 4) * 2 - 1, torch.rand(1, 16, 16, 4) * 2 - 1, torch.rand(1, 32, 32, 4) * 2 - 1]
    g0 = densenet.generator(1, init=True, device="cpu", u=u, **kw)
    assert g0.shape == (1, 32, 32, 3)
    perturb(densenet.discriminator, 3); perturb(densenet.generator, 4)
    f = densenet.discriminator(x, **kw).detach().double().numpy()
    ref = no.densenet_discriminator(x.double().numpy(), table_of(densenet.discriminator), 3, 4)
    assert f.shape == ref.shape and np.abs(f - ref).max() / np.abs(ref).max() < 2e-5
    g = densenet.generator(1, u=u, **kw).detach().double().numpy()
    gref = no.densenet_generator([t.double().numpy() for t in u], table_of(densenet.generator), 3, 4)
    assert np.abs(g - gref).max() < 2e-5
    densenet.discriminator.reset(); densenet.generator.reset()


def test_densenet_default_shapes():
    densenet.discriminator.reset(); densenet.generator.reset()
    f = densenet.discriminator(torch.zeros(1, 32, 32, 3) + 0.1, init=True, device="cpu")
    g = densenet.generator(1, init=True, device="cpu")
    assert f.shape == (1, 7296) and g.shape == (1, 32, 32, 3)
    assert densenet.discriminator.store.num_params() == 7453016 and densenet.generator.store.num_params() == 6012422
    densenet.discriminator.reset(); densenet.generator.reset()


def test_gradients_flow_to_flat_buffers_and_match_directional_derivative():
    """grad_ys-driven backward (train.py:112): <autograd grad, delta> == d/d eps of <grad_ys, f(theta + eps delta)> (oracle, fp64)."""
    dcgan.discriminator.reset()
    torch.manual_seed(2)
    x = torch.rand(1, 32, 32, 3) * 2 - 1
    dcgan.discriminator(x, init=True, device="cpu")
    perturb(dcgan.discriminator, 5)
    f = dcgan.discriminator(x)
    gy = torch.randn_like(f)
    (grad,) = torch.autograd.grad([f], [dcgan.discriminator.flat], grad_outputs=[gy])
    delta = torch.randn_like(grad) * 1e-3
    table = table_of(dcgan.discriminator)
    def fwd(eps):
        t, off = {}, 0
        for n, p in dcgan.discriminator.named_parameters():
            k = p.numel()
            t[n] = table[n] + eps * delta[off:off + k].reshape(p.shape).double().numpy()
            off += k
        return no.dcgan_discriminator(x.double().numpy(), t)
    eps = 1e-5        # the net is piecewise linear (CReLU): keep the probe far smaller than the distance to the next kink
    fd = np.sum(gy.double().numpy() * (fwd(eps) - fwd(-eps))) / (2 * eps)
    an = float((grad.double() * delta.double()).sum())
    assert abs(fd - an) <= 1e-3 * max(abs(fd), abs(an)) + 1e-9
    dcgan.discriminator.reset()


def test_adam_oracle_formula():
    p, g = np.array([1.0, -2.0]), np.array([0.5, -0.25])
    p1, v1, mg1 = no.adam_step(p, g, np.zeros(2), np.zeros(2), 1, 3e-4, 0.5, 0.999)
    # t = 1: v_hat = g, mg_hat = g^2  ->  step = lr * g / sqrt(g^2 + 1e-8)
    np.testing.assert_allclose(p1, p - 3e-4 * g / np.sqrt(g * g + 1e-8), rtol=1e-12)


def test_dcgan_image_size_extension_shapes():
    """BASELINE config 5 (64 x 64 synthetic images; not in the reference, whose shapes are hard-wired to 32 x 32): the
    generator seeded at 8 x 8 emits [B, 64, 64, 3]; the fully convolutional critic maps it to 8*8*2048 = 131072 unit-norm
    features; image_size = 32 leaves the reference's parameter counts untouched."""
    dcgan.discriminator.reset(); dcgan.generator.reset()
    torch.manual_seed(0)
    with torch.no_grad():
        x = dcgan.generator(2, init=True, device="cpu", image_size=64)
        assert tuple(x.shape) == (2, 64, 64, 3) and float(x.abs().max()) <= 1.0
        f = dcgan.discriminator(x, init=True, device="cpu")
    assert tuple(f.shape) == (2, 131072)
    np.testing.assert_allclose(f.pow(2).sum(1).numpy(), 1.0, rtol=1e-5)
    assert dcgan.discriminator.store.num_params() == 34419840                  # the critic has no size-dependent variable
    assert dcgan.generator.store.num_params() == 37761926 + (4 - 1) * (100 + 2) * 2 * 4 * 4 * 1024     # only the seed dense layer grows
    dcgan.discriminator.reset(); dcgan.generator.reset()
