"""Sample-grid writer with the reference's tiling (TG/tflib/save_images.py:9-37): a batch `X` -- [n, C, H, W],
[n, H, W] or flat [n, H*W] -- becomes one image of `rows x n/rows` tiles, rows = the largest divisor of n that is
<= sqrt(n); float input is scaled by 255.99 and truncated to uint8 (:11-12).  Written with Pillow (`scipy.misc.imsave`,
which the reference uses, no longer exists)."""
import numpy as np


def tile_images(X):
    """The grid the reference assembles before `imsave`: float64 array [h*rows, w*cols(, 3)]."""
    X = np.asarray(X)
    if isinstance(X.flatten()[0], np.floating):
        X = (255.99 * X).astype('uint8')
    n_samples = X.shape[0]
    rows = int(np.sqrt(n_samples))
    while n_samples % rows != 0:
        rows -= 1
    nh, nw = rows, n_samples // rows
    if X.ndim == 2:
        side = int(np.sqrt(X.shape[1]))
        X = np.reshape(X, (X.shape[0], side, side))
    if X.ndim == 4:
        X = X.transpose(0, 2, 3, 1)                     # BCHW -> BHWC
        h, w = X[0].shape[:2]
        img = np.zeros((h * nh, w * nw, 3))
    elif X.ndim == 3:
        h, w = X[0].shape[:2]
        img = np.zeros((h * nh, w * nw))
    else:
        raise ValueError('save_images: X must be 2-, 3- or 4-dimensional')
    for n, x in enumerate(X):
        j, i = divmod(n, nw)
        img[j * h:j * h + h, i * w:i * w + w] = x
    return img


def _to_uint8(img):
    """scipy.misc.imsave semantics (bytescale): the grid's own min..max is stretched to 0..255."""
    lo, hi = float(img.min()), float(img.max())
    if hi == lo:
        return np.zeros(img.shape, dtype='uint8')
    return ((img - lo) * (255.0 / (hi - lo)) + 0.5).clip(0, 255).astype('uint8')


def save_images(X, save_path):
    from PIL import Image
    if hasattr(X, 'detach'):                            # torch tensor (device or host)
        X = X.detach().cpu().numpy()
    img = _to_uint8(tile_images(X))
    Image.fromarray(img, 'RGB' if img.ndim == 3 else 'L').save(save_path)
