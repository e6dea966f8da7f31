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
    open(os.path.join(tmp_path, "med_gan_params-2399.index"), "wb").write(b"\0")         # looks like a TF checkpoint prefix
    with pytest.raises(ImportError):
        checkpoint.load_variables(os.path.join(tmp_path, "med_gan_params-2399"))         # ... and there is no tensorflow here
    with pytest.raises(FileNotFoundError):
        checkpoint.load_variables(os.path.join(tmp_path, "nothing-here"))
    # the reference's extensionless torch archive (what Trainer.save writes as `med_gan_params-<epoch>`) is sniffed, not guessed
    ext_less = os.path.join(tmp_path, "med_gan_params-199")
    torch.save({"discriminator": {n: p.detach().clone() for n, p in dcgan.discriminator.named_parameters()},
                "generator": {n: p.detach().clone() for n, p in dcgan.generator.named_parameters()}}, ext_less)
    again = checkpoint.load_variables(ext_less)
    assert sorted(again) == names and np.array_equal(again["generator/dense_0/g"], saved["generator/dense_0/g"])
    dcgan.discriminator.reset(); dcgan.generator.reset()


def test_reference_style_variable_dump_loads_and_reproduces_the_reference_features(tmp_path):
    """a `med_gan_params-*` style dump keyed by TensorFlow variable names (`<scope>/<layer>/V:0`, HWIO kernels,
    [in, out] dense matrices) -- written here from the same name-seeded variables the reference graph was evaluated with in
    tests/golden/make_reference_golden.py -- goes through checkpoint.load_variables / assign and must reproduce the features
    and the generated image of the REFERENCE's own models/dcgan.py (tests/golden/ref_model_dcgan.npz)."""
    from otgan_b200.models import dcgan
    from otgan_b200.utils import checkpoint
    from tests.test_reference_golden import seeded_variable
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_model_dcgan.npz"))
    dcgan.discriminator.reset(); dcgan.generator.reset()
    with torch.no_grad():
        dcgan.discriminator(torch.zeros(2, 32, 32, 3) + 0.1, init=True, device="cpu")
        dcgan.generator(2, init=True, device="cpu")
    tpls = (dcgan.discriminator, dcgan.generator)
    dump = {n + ":0": seeded_variable(n, tuple(p.shape)) for t in tpls for n, p in t.named_parameters()}
    path = os.path.join(tmp_path, "med_gan_params-2399.npz")
    np.savez(path, **dump)
    checkpoint.assign(tpls, checkpoint.load_variables(os.path.join(tmp_path, "med_gan_params-2399")))    # prefix form
    with torch.no_grad():
        f = dcgan.discriminator(torch.from_numpy(g["x"])).double().numpy()
        img = dcgan.generator(2, u=torch.from_numpy(g["u0"])).double().numpy()
    assert np.abs(f - g["features"]).max() / np.abs(g["features"]).max() < 2e-5
    assert np.abs(img - g["image"]).max() < 2e-5
    dcgan.discriminator.reset(); dcgan.generator.reset()


def test_inception_score_hook_arithmetic_and_contract():
    """utils/inception.py:43-52: the split / KL / exp arithmetic against a direct restatement, the reference's input checks, and
    the loud failure without a classifier."""
    import numpy as np
    import pytest
    from otgan_b200.utils import inception
    rng = np.random.RandomState(0)
    imgs = [rng.randint(0, 256, size=(32, 32, 3)).astype(np.float64) for _ in range(250)]
    W = rng.randn(3 * 32 * 32, 7) * 1e-3

    def predict(batch):
        z = batch.reshape(len(batch), -1) @ W
        z = np.exp(z - z.max(1, keepdims=True))
        return z / z.sum(1, keepdims=True)

    mean, std = inception.get_inception_score(imgs, splits=5, predict_fn=predict)
    preds = np.concatenate([predict(np.stack(imgs[i:i + 100]).astype(np.float32)) for i in range(0, 250, 100)])
    ref = []
    for part in np.array_split(preds, 5):
        py = part.mean(0, keepdims=True)
        ref.append(np.exp(np.mean(np.sum(part * (np.log(part) - np.log(py)), 1))))
    assert abs(mean - np.mean(ref)) < 1e-12 and abs(std - np.std(ref)) < 1e-12 and 1.0 <= mean <= 7.0
    with pytest.raises(RuntimeError):
        inception.get_inception_score(imgs)
    with pytest.raises(AssertionError):
        inception.get_inception_score([im / 255.0 for im in imgs], predict_fn=predict)      # reference check: values in [0, 255]
