"""DenseNet critic / generator of the reference (models/densenet.py), re-hosted on otgan_b200.utils.nn.

    discriminator(x, init=False, layers_per_block=16, filters_per_layer=16, nonlinearity='crelu', ema=None) -> [B, 7296]   :7-47
    generator(batch_size, init=False, layers_per_block=16, filters_per_layer=16, nonlinearity='crelu', ema=None)            :51-90
Dense blocks pass Python LISTS of tensors to nn.conv2d; CReLU interleaves per list element (utils/nn.py:198-200).
"""
import numpy as np
import torch

from ..utils import nn
from ..utils.nn import arg_scope


# //// discriminator ////
def disc_spec(x, init=False, layers_per_block=16, filters_per_layer=16, nonlinearity='crelu', ema=None, **kwargs):
    with arg_scope([nn.conv2d, nn.dense, nn.dense_block], counters={}, init=init, weight_norm=True, ema=ema):

        def block(x):
            if type(x) is not list:
                x = [x]
            return nn.dense_block(x, layers_per_block, filters_per_layer, pre_activation=nonlinearity)      # :10-15

        def downsample(x):
            if isinstance(x, nn.Crelu8Tensor):
                return nn.conv2d(x, x.raw_channels // 2, pre_activation=nonlinearity, stride=[2, 2])
            if type(x) is not list:
                x = [x]
            return nn.conv2d(x, int(np.sum([int(xi.shape[-1]) for xi in x])) // 2, pre_activation=nonlinearity, stride=[2, 2])

        x = nn.conv2d(x, 2 * filters_per_layer, pre_activation=None)
        x = block(x)
        x = downsample(x)
        x = block(x)
        x = downsample(x)
        x = block(x)
        x = downsample(x)

        x = torch.cat(x, 3) if isinstance(x, list) else x
        # :38-42  concat([relu(x), relu(-x)], 3) -> reshape [B, -1] -> x / sqrt(sum(x^2))
        return nn.crelu_l2norm(x)


discriminator = nn.make_template('discriminator', disc_spec)


# //// generator ////
def gen_spec(batch_size, init=False, layers_per_block=16, filters_per_layer=16, nonlinearity='crelu', ema=None, u=None, **kwargs):
    device = nn._tls.store.device
    if u is None:
        u = [torch.rand((batch_size, 100), device=device) * 2.0 - 1.0,
             torch.rand((batch_size, 8, 8, filters_per_layer), device=device) * 2.0 - 1.0,
             torch.rand((batch_size, 16, 16, filters_per_layer), device=device) * 2.0 - 1.0,
             torch.rand((batch_size, 32, 32, filters_per_layer), device=device) * 2.0 - 1.0]

    with arg_scope([nn.conv2d, nn.dense, nn.dense_block], counters={}, init=init, weight_norm=True, ema=ema):

        def block(x):
            if type(x) is not list:
                x = [x]
            return nn.dense_block(x, layers_per_block, filters_per_layer, pre_activation=nonlinearity)      # :56-61

        def upsample(x):
            if isinstance(x, nn.Crelu8Tensor):           # resize commutes with the element-wise CReLU
                return nn.conv2d(x.upsample2x(), x.raw_channels // 2, pre_activation=nonlinearity)
            if type(x) is list:
                x = torch.cat(x, 3)
            xs = list(x.shape)
            x = nn.resize_nearest_neighbor(x, [xs[1] * 2, xs[2] * 2])
            x = nn.conv2d(x, xs[3] // 2, pre_activation=nonlinearity)
            return x

        x = nn.dense(u[0], 8 * 8 * filters_per_layer, pre_activation=None)
        x = [x.reshape(batch_size, 8, 8, filters_per_layer), u[1]]
        x = block(x)
        x = upsample(x)
        x = [x, u[2]]
        x = block(x)
        x = upsample(x)
        x = [x, u[3]]
        x = block(x)
        x = torch.tanh(nn.conv2d(x, 3, pre_activation=nonlinearity, init_scale=0.1))
        return x


generator = nn.make_template('generator', gen_spec)
