"""Loss-dictionary templates with the reference's exact keys (reference utils.py:32-53), the metrics dump of the
run drivers (pix2pix.py:436-440, cycle_gan.py:477-481) and the sample-image panels that stand in for the
matplotlib figures of ``generate_images`` (pix2pix.py:220-246, cycle_gan.py:179-204; matplotlib is not a
dependency here: the panels are written with PIL, pixel values mapped ``x*0.5+0.5`` exactly as the reference plots)."""
import json
import os

import numpy as np



def pix2pix_losses():
    return {'Generator Total Loss': [],
            'Generator Loss (Primary)': [],
            'Generator Loss (Secondary)': [],
            'Discriminator Loss': []}


def cyclegan_losses():
    return {'X->Y Generator Loss': [],
            'Y->X Generator Loss': [],
            'Total Cycle Loss': [],
            'Total X->Y Generator Loss': [],
            'Total Y->X Generator Loss': [],
            'Discriminator X Loss': [],
            'Discriminator Y Loss': []}


def dump_metrics(log_dir: str, train_metrics: dict, val_metrics: dict):
    """Reference pix2pix.py:436-440 / cycle_gan.py:477-481: ``train_metrics.json`` and ``val_metrics.json`` in the log
    directory, the dicts ``fit`` returns (same keys, one mean loss per epoch)."""
    os.makedirs(log_dir, exist_ok=True)
    paths = []
    for name, d in (("train_metrics.json", train_metrics), ("val_metrics.json", val_metrics)):
        path = os.path.join(log_dir, name)
        with open(path, "w") as f:
            json.dump(d, f)
        paths.append(path)
    return paths


def save_panel(path_filename: str, images, channels: int):
    """Side-by-side panel of (H, W, C) float images in [-1, 1] ('Input Image', ['Ground Truth',] 'Predicted Image'
    in the reference's figures), written as PNG; values are mapped x*0.5+0.5 like the reference's imshow calls."""
    from PIL import Image
    tiles = []
    for im in images:
        a = np.clip(np.asarray(im, dtype=np.float32) * 0.5 + 0.5, 0.0, 1.0)
        a = (a * 255.0 + 0.5).astype(np.uint8)
        if a.ndim == 2:
            a = a[:, :, None]
        if a.shape[2] == 1:
            a = np.repeat(a, 3, axis=2)
        tiles.append(a[:, :, :3])
    gap = np.full((tiles[0].shape[0], 8, 3), 255, dtype=np.uint8)
    row = [tiles[0]]
    for t in tiles[1:]:
        row += [gap, t]
    os.makedirs(os.path.dirname(os.path.abspath(path_filename)), exist_ok=True)
    Image.fromarray(np.concatenate(row, axis=1)).save(path_filename)
    return path_filename


def _gauss_kernel(size: int, sigma: float):
    x = np.arange(size, dtype=np.float64) - (size - 1) / 2.0
    g = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return g / g.sum()


def ssim(img1, img2, max_val=255, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03):
    """``tf.image.ssim`` on host arrays (B, H, W, C): Gaussian-weighted local statistics over 'valid' windows,
    SSIM map = luminance * contrast-structure, mean over the window positions and channels -> one value per image.
    Used by the reference's non-default generator loss (pix2pix.py:182-184), which evaluates it on the INPUT and
    the TARGET (both constants of the step)."""
    a = np.asarray(img1, dtype=np.float64); b = np.asarray(img2, dtype=np.float64)
    g = _gauss_kernel(filter_size, filter_sigma)
    c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2

    def blur(x):                                          # separable 'valid' correlation along H then W
        h = np.lib.stride_tricks.sliding_window_view(x, filter_size, axis=1) @ g
        return np.lib.stride_tricks.sliding_window_view(h, filter_size, axis=2) @ g
    mu1, mu2 = blur(a), blur(b)
    s11, s22, s12 = blur(a * a) - mu1 * mu1, blur(b * b) - mu2 * mu2, blur(a * b) - mu1 * mu2
    lum = (2 * mu1 * mu2 + c1) / (mu1 * mu1 + mu2 * mu2 + c1)
    cs = (2 * s12 + c2) / (s11 + s22 + c2)
    return (lum * cs).mean(axis=(1, 2, 3))
