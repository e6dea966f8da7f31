"""DCGAN critic / generator of the reference (models/dcgan.py), re-hosted on otgan_b200.utils.nn.

Same constructor signatures and variable layout as /root/reference/models/dcgan.py:
    discriminator(x, init=False, nonlinearity='crelu', ema=None, **kwargs) -> [B, 32768] L2-normalised features   :7-24
    generator(batch_size, init=False, nonlinearity='crelu', ema=None, **kwargs) -> [B, 32, 32, 3]                  :28-54
`x` is NHWC like the reference; both are make_template callables that share their variables across calls.
"""
import numpy as np
import torch

from ..utils import nn
from ..utils.nn import arg_scope


# //// discriminator ////
def disc_spec(x, init=False, nonlinearity='crelu', ema=None, **kwargs):
    with arg_scope([nn.conv2d, nn.dense], counters={}, init=init, weight_norm=True, ema=ema):
        x = nn.conv2d(x, 128, filter_size=[5, 5], pre_activation=None)
        x = nn.conv2d(x, 256, filter_size=[5, 5], pre_activation=nonlinearity, stride=[2, 2])
        x = nn.conv2d(x, 512, filter_size=[5, 5], pre_activation=nonlinearity, stride=[2, 2])
        x = nn.conv2d(x, 1024, filter_size=[5, 5], pre_activation=nonlinearity, stride=[2, 2])

        # :16-19  concat([relu(x), relu(-x)], 3) -> reshape [B, -1] (NHWC order) -> x / sqrt(sum(x^2)) (no epsilon)
        return nn.crelu_l2norm(x)


discriminator = nn.make_template('discriminator', disc_spec)


# //// generator ////
def gen_spec(batch_size, init=False, nonlinearity='crelu', ema=None, u=None, **kwargs):
    device = nn._tls.store.device
    if u is None:
        u = torch.rand((batch_size, 100), device=device) * 2.0 - 1.0            # tf.random_uniform(-1, 1)  :30
    with arg_scope([nn.conv2d, nn.dense], counters={}, init=init, weight_norm=True, ema=ema):
        x = nn.dense(u, 2 * 4 * 4 * 1024, pre_activation=None)
        x, l = torch.chunk(x, 2, 1)
        x = x * torch.sigmoid(l)                                                # gated linear unit  :35-36
        x = x.reshape(batch_size, 4, 4, 1024)
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
