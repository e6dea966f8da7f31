"""CPU checks (numpy, float64) of the algebra behind the fused upsample + convolution kernels (conv_tc.cu, make_submap /
up2_presum / up2_unsum) and the 3-channel-layer decomposition (conv_narrow.cu):

  conv_SAME(resize_nearest_neighbor(x, 2x), W)  ==  for each output parity (a, b): conv_3x3(x, W'_ab) written to (2i+a, 2j+b)
  with W'_ab[ri, ci] = sum of the W[kh, kw] whose tap lands on low-resolution offset (dmin_a + ri, dmin_b + ci),
  floor((a + kh - pad) / 2) being the low-resolution row a tap reads (models/dcgan.py:37-46, utils/nn.py:234-241).

These restate the host-side tables of the CUDA path in plain numpy so that the identity itself is pinned without a GPU."""
import numpy as np
import pytest


def submap(k, pad):
    """Python restatement of make_submap (conv_tc.cu): per output parity a, the low-res offset of every filter tap."""
    dmin, idx = [], []
    for a in range(2):
        offs = [(a + kk - pad) // 2 for kk in range(k)]          # python // is floor division
        dmin.append(offs[0])
        idx.append([o - offs[0] for o in offs])
    n1 = [i[-1] + 1 for i in idx]
    assert n1[0] == n1[1]
    return dmin, idx, n1[0]


def conv_same(x, w, pad):
    """x: [H, W, Cin], w: [k, k, Cin, Cout] -> [H, W, Cout], zero padding `pad` on every side (stride 1)."""
    H, W, _ = x.shape
    k = w.shape[0]
    xp = np.zeros((H + 2 * pad, W + 2 * pad, x.shape[2]))
    xp[pad:pad + H, pad:pad + W] = x
    y = np.zeros((H, W, w.shape[3]))
    for a in range(k):
        for b in range(k):
            y += xp[a:a + H, b:b + W] @ w[a, b]
    return y


def presum(w, pad):
    k = w.shape[0]
    dmin, idx, n1 = submap(k, pad)
    sub = np.zeros((2, 2, n1, n1) + w.shape[2:])
    for a in range(2):
        for b in range(2):
            for kh in range(k):
                for kw in range(k):
                    sub[a, b, idx[a][kh], idx[b][kw]] += w[kh, kw]
    return sub, dmin, idx, n1


@pytest.mark.parametrize("k", [3, 5])
def test_upsample_then_conv_equals_four_subfilter_convs(k):
    rng = np.random.RandomState(k)
    pad = (k - 1) // 2
    x = rng.randn(6, 4, 3)
    w = rng.randn(k, k, 3, 5)
    ref = conv_same(x.repeat(2, 0).repeat(2, 1), w, pad)
    sub, dmin, idx, n1 = presum(w, pad)
    assert n1 == {3: 2, 5: 3}[k]                                   # 9 instead of 25 taps for the generator's 5x5 layers
    H, W, _ = x.shape
    out = np.zeros_like(ref)
    for a in range(2):
        for b in range(2):
            for ri in range(n1):
                for ci in range(n1):
                    dh, dw = dmin[a] + ri, dmin[b] + ci
                    for i in range(H):
                        for j in range(W):
                            if 0 <= i + dh < H and 0 <= j + dw < W:   # zero outside the LOW-resolution image (TMA zero fill)
                                out[2 * i + a, 2 * j + b] += x[i + dh, j + dw] @ sub[a, b, ri, ci]
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-12)


@pytest.mark.parametrize("k", [3, 5])
def test_unsum_is_the_chain_rule_of_presum(k):
    """dL/dW[kh, kw] = sum over parities of dL/dW'_ab[idx_a[kh], idx_b[kw]] (up2_unsum_kernel): checked through the
    linear map W -> W' with a random cotangent."""
    rng = np.random.RandomState(10 + k)
    pad = (k - 1) // 2
    w = rng.randn(k, k, 2, 3)
    sub, dmin, idx, n1 = presum(w, pad)
    g_sub = rng.randn(*sub.shape)                                  # cotangent of W'
    g_w = np.zeros_like(w)
    for kh in range(k):
        for kw in range(k):
            for a in range(2):
                for b in range(2):
                    g_w[kh, kw] += g_sub[a, b, idx[a][kh], idx[b][kw]]
    dw = rng.randn(*w.shape)                                       # <g_sub, presum(dw)> == <unsum(g_sub), dw>
    lhs = float((g_sub * presum(dw, pad)[0]).sum())
    rhs = float((g_w * dw).sum())
    assert abs(lhs - rhs) < 1e-10 * max(1.0, abs(lhs))


def test_three_channel_layer_as_one_gemm_plus_shift():
    """conv_narrow.cu: a convolution with 3 OUTPUT channels == one [pixels, Cin] x [Cin, 75] GEMM (z) followed by
    y[px, c] = sum_t z[px + off_t, t*3 + c]; with 3 INPUT channels the filter gradient == wide^T . im2col(narrow)."""
    rng = np.random.RandomState(3)
    k, pad, H, W, Cin, Cout = 5, 2, 6, 5, 7, 3
    x = rng.randn(H, W, Cin)
    w = rng.randn(k, k, Cin, Cout)
    ref = conv_same(x, w, pad)
    z = x.reshape(-1, Cin) @ w.transpose(2, 0, 1, 3).reshape(Cin, k * k * Cout)          # z[px, t*Cout + c]
    z = z.reshape(H, W, k * k, Cout)
    y = np.zeros_like(ref)
    for a in range(k):
        for b in range(k):
            dh, dw = a - pad, b - pad
            for i in range(H):
                for j in range(W):
                    if 0 <= i + dh < H and 0 <= j + dw < W:
                        y[i, j] += z[i + dh, j + dw, a * k + b]
    np.testing.assert_allclose(y, ref, rtol=0, atol=1e-12)
    # filter gradient of a 3-input-channel layer: dW[t*3 + ci, co] = sum_px col[px, t*3 + ci] * dy[px, co]
    x3 = rng.randn(H, W, 3)
    dy = rng.randn(H, W, 4)
    col = np.zeros((H, W, k * k * 3))
    for a in range(k):
        for b in range(k):
            for i in range(H):
                for j in range(W):
                    if 0 <= i + a - pad < H and 0 <= j + b - pad < W:
                        col[i, j, (a * k + b) * 3:(a * k + b) * 3 + 3] = x3[i + a - pad, j + b - pad]
    dw = (col.reshape(-1, k * k * 3).T @ dy.reshape(-1, 4)).reshape(k, k, 3, 4)
    eps = 1e-6
    w3 = rng.randn(k, k, 3, 4)
    for (a, b, ci, co) in [(0, 0, 0, 0), (2, 3, 1, 2), (4, 4, 2, 3)]:
        wp = w3.copy(); wp[a, b, ci, co] += eps
        num = ((conv_same(x3, wp, pad) - conv_same(x3, w3, pad)) * dy).sum() / eps
        assert abs(num - dw[a, b, ci, co]) < 1e-5
