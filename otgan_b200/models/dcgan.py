"""DCGAN critic / generator of the reference (models/dcgan.py), re-hosted on otgan_b200.utils.nn.

Same constructor signatures and variable layout as /root/reference/models/dcgan.py:
    discriminator(x, init=False, nonlinearity='crelu', ema=None, **kwargs) -> [B, 32768] L2-normalised features   :7-24
    generator(batch_size, init=False, nonlinearity='crelu', ema=None, **kwargs) -> [B, 32, 32, 3]                  :28-54
`x` is NHWC like the reference; both are make_template callables that share their variables across calls.

Extension (not in the reference, whose shapes are hard-wired to 32 x 32: train.py:52,67; BASELINE config 5 asks for 64 x 64):
`generator(..., image_size=S)` seeds the same three resize + convolution stages at S/8 x S/8 (dense layer of
2 * (S/8)^2 * 1024 units) and emits [B, S, S, 3]; the critic is fully convolutional, so a [B, S, S, 3] input gives
(S/8)^2 * 2048 features (131 072 at S = 64).  S = 32 is exactly the reference.
"""
import torch

from ..utils import nn
from ..utils.nn import arg_scope


# //// discriminator ////
def disc_spec(x, init=False, nonlinearity='crelu', ema=None, **kwargs):
    with arg_scope([nn.conv2d, nn.dense], counters={}, init=init, weight_norm=True, ema=ema):
        # crelu_out: the CReLU that the NEXT layer applies to its input (pre_activation) is written by THIS layer's convolution
        # epilogue on the tcgen05 path (nn.CreluOut); same values, one pass over the activations less per layer
        fuse = nonlinearity == 'crelu'
        x = nn.conv2d(x, 128, filter_size=[5, 5], pre_activation=None, crelu_out=fuse)
        x = nn.conv2d(x, 256, filter_size=[5, 5], pre_activation=nonlinearity, stride=[2, 2], crelu_out=fuse)
        x = nn.conv2d(x, 512, filter_size=[5, 5], pre_activation=nonlinearity, stride=[2, 2], crelu_out=fuse)
        x = nn.conv2d(x, 1024, filter_size=[5, 5], pre_activation=nonlinearity, stride=[2, 2])

        # :16-19  concat([relu(x), relu(-x)], 3) -> reshape [B, -1] (NHWC order) -> x / sqrt(sum(x^2)) (no epsilon)
        return nn.crelu_l2norm(x)


discriminator = nn.make_template('discriminator', disc_spec)


# //// generator ////
def gen_spec(batch_size, init=False, nonlinearity='crelu', ema=None, u=None, image_size=32, **kwargs):
    device = nn._tls.store.device
    s0 = image_size // 8                      # side of the seed feature map: 4 for the reference's 32 x 32 images
    assert image_size == 8 * s0 and s0 >= 1, "image_size must be a multiple of 8"
    if u is None:
        u = torch.rand((batch_size, 100), device=device) * 2.0 - 1.0            # tf.random_uniform(-1, 1)  :30
    with arg_scope([nn.conv2d, nn.dense], counters={}, init=init, weight_norm=True, ema=ema):
        x = nn.dense(u, 2 * s0 * s0 * 1024, pre_activation=None)
        # x, l = split(x, 2, 1); x *= sigmoid(l)  (gated linear unit, :35-36) -- the same split as nn.glu's channel split on a
        # [B, 1, 1, 2F] view: one fused kernel forward / backward on the GPU path
        x = nn.glu(x.reshape(batch_size, 1, 1, 2 * s0 * s0 * 1024))
        x = x.reshape(batch_size, s0, s0, 1024)
        x = nn.upsample2x(x)              # tf.image.resize_nearest_neighbor(x, [8, 8])  :37-38 (fused into the next conv2d)
        x = nn.conv2d(x, 2 * 512, filter_size=[5, 5], pre_activation=None)
        x = nn.glu(x, upsample=True)      # x, l = split(x, 2, 3); x *= sigmoid(l); resize_nearest_neighbor(x, [16, 16])  :39-42
        x = nn.conv2d(x, 2 * 256, filter_size=[5, 5], pre_activation=None)
        x = nn.glu(x, upsample=True)      # ... resize_nearest_neighbor(x, [32, 32])                                        :43-46
        x = nn.conv2d(x, 2 * 128, filter_size=[5, 5], pre_activation=None)
        x = nn.glu(x)                     # x, l = split(x, 2, 3); x *= sigmoid(l)                                          :47-48
        x = torch.tanh(nn.conv2d(x, 3, filter_size=[5, 5], pre_activation=None, init_scale=0.1))
        return x


generator = nn.make_template('generator', gen_spec)
