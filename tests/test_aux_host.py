"""CPU tests of the side utilities kept for drop-in use of train.py: sample tiling (utils/plotting.py:9-74) and parameter
import / export under the TensorFlow variable names (train.py:60,190-193,276)."""
import os

import numpy as np
import pytest
import torch


def test_img_tile_layout_and_png(tmp_path):
    from otgan_b200.utils import plotting
    imgs = np.stack([np.full((4, 6, 3), v, dtype=np.float32) for v in np.linspace(-1, 1, 7)])
    t = plotting.img_tile(imgs, aspect_ratio=1.0, border_color=1.0, stretch=False)
    rows, cols = plotting.grid_shape_for(7, (4, 6), 1.0)
    assert rows * cols >= 7 and t.shape == ((4 + 1) * rows - 1, (6 + 1) * cols - 1, 3)
    assert np.all(t[0:4, 0:6] == imgs[0]) and np.all(t[0:4, 7:13] == imgs[1])          # row-major placement
    assert np.all(t[4, :] == 1.0) and np.all(t[:4, 6] == 1.0)                             # border colour
    t2 = plotting.img_tile(imgs, tile_shape=(2, 2), border=0)
    assert t2.shape == (8, 12, 3) and np.all(t2[4:, 6:] == imgs[3])
    with pytest.raises(ValueError):
        plotting.img_tile(np.zeros((2, 2)))
    path = os.path.join(tmp_path, "tile.png")
    plotting.save_tile_img(t, path)
    from PIL import Image
    im = np.asarray(Image.open(path))
    assert im.shape == t.shape and im.dtype == np.uint8 and im[0, 0, 0] == 0 and im[4, 0, 0] == 255
    s = plotting.img_stretch(np.array([2.0, 4.0, 6.0]))
    assert s[0] == 0.0 and abs(s[2] - 1.0) < 1e-9


def test_checkpoint_round_trip_by_tensorflow_variable_names(tmp_path):
    from otgan_b200.models import dcgan
    from otgan_b200.utils import checkpoint
    dcgan.discriminator.reset(); dcgan.generator.reset()
    torch.manual_seed(0)
    with torch.no_grad():
        dcgan.discriminator(torch.zeros(2, 32, 32, 3), init=True, device="cpu")
        dcgan.generator(2, init=True, device="cpu")
    tpls = (dcgan.discriminator, dcgan.generator)
    path = os.path.join(tmp_path, "med_gan_params-7.npz")
    names = checkpoint.export_npz(tpls, path)
    assert "discriminator/conv2d_1/V" in names and "generator/dense_0/g" in names and "generator/conv2d_3/b" in names
    saved = checkpoint.load_variables(path)
    assert saved["discriminator/conv2d_1/V"].shape == (5, 5, 256, 256)                   # HWIO, as tf.nn.conv2d stores it
    before = [t.flat.detach().clone() for t in tpls]
    v0 = [t.store.version for t in tpls]
    with torch.no_grad():
        for t in tpls:
            t.flat.add_(1.0)
    assigned = checkpoint.assign(tpls, saved)
    assert sorted(assigned) == names
    for t, b, v in zip(tpls, before, v0):
        assert torch.equal(t.flat.detach()[: t.store.num_params()], b[: t.store.num_params()]) and t.store.version == v + 1
    bad = dict(saved)
    bad["generator/conv2d_0/V"] = np.zeros((3, 3, 1, 1), np.float32)
    with pytest.raises(ValueError):
        checkpoint.assign(tpls, bad)
    del bad["generator/conv2d_0/V"]
    with pytest.raises(KeyError):
        checkpoint.assign(tpls, bad)
    with pytest.raises(ImportError):
        checkpoint.load_variables(os.path.join(tmp_path, "med_gan_params-2399"))         # TF checkpoint prefix, no tensorflow here
    dcgan.discriminator.reset(); dcgan.generator.reset()
