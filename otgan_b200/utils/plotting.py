"""PIL + numpy sample tiling: the two helpers of the reference's utils/plotting.py that train.py actually calls
(img_tile :29-74, save_tile_img :9-13, used at train.py:233-243).  matplotlib is not needed for either; the rest of the
reference module (filter tiles, raster plots) has no caller in the tree and is out of scope."""
import numpy as np


def img_stretch(img):
    """Rescale to [0, 1] (utils/plotting.py:23-27)."""
    img = np.asarray(img, dtype=float)
    lo = img.min()
    return (img - lo) / (img.max() - lo + 1e-12)


def grid_shape_for(n_imgs, img_hw, aspect_ratio=1.0):
    """Rows x columns of a near-square grid for n_imgs images of size img_hw (utils/plotting.py:45-50)."""
    ar = aspect_ratio * img_hw[1] / float(img_hw[0])
    return int(np.ceil(np.sqrt(n_imgs * ar))), int(np.ceil(np.sqrt(n_imgs / ar)))


def img_tile(imgs, aspect_ratio=1.0, tile_shape=None, border=1, border_color=0, stretch=False):
    """Tile [n, H, W] or [n, H, W, C] images into one grid image, row-major, `border` pixels of `border_color` between
    tiles; with tile_shape only rows*cols images are used.  Same arguments and result as the reference function."""
    imgs = np.asarray(img_stretch(imgs) if stretch else imgs)
    if imgs.ndim not in (3, 4):
        raise ValueError('imgs has wrong number of dimensions.')
    n, h, w = imgs.shape[:3]
    if tile_shape is None:
        rows, cols = grid_shape_for(n, (h, w), aspect_ratio)
    else:
        assert len(tile_shape) == 2
        rows, cols = int(tile_shape[0]), int(tile_shape[1])
    out = np.full(((h + border) * rows - border, (w + border) * cols - border) + imgs.shape[3:], border_color, dtype=float)
    for idx in range(min(n, rows * cols)):
        i, j = divmod(idx, cols)
        y0, x0 = (h + border) * i, (w + border) * j
        out[y0:y0 + h, x0:x0 + w] = imgs[idx]
    return out


def save_tile_img(imgs, path):
    """[-1, 1] float image -> 8-bit PNG (utils/plotting.py:9-13)."""
    from PIL import Image
    Image.fromarray(((np.asarray(imgs) + 1.0) * 127.5).astype(np.uint8)).save(path)
