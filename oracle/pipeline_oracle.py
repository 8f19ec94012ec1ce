"""CPU restatement of the reference's per-image input pipeline.  TEST INFRASTRUCTURE ONLY: imported
by tests/ (and nothing else); the product path (gan_b200/) never touches it.

PARITY UNPINNED: the reference runs these steps through TensorFlow 2.6 (`tf.image.resize`,
`tf.image.random_crop`, `tf.image.flip_left_right`), which cannot be installed here, and ships no
tests or fixtures for them.  The one non-obvious piece of arithmetic, the nearest-neighbour source
index, restates TF2's published kernel (ResizeNearestNeighbor with half_pixel_centers=True, which is
what `tf.image.resize(..., method=NEAREST_NEIGHBOR)` calls in TF >= 2.0): float32
`min(floor((o + 0.5) * (in / out)), in - 1)`.  Everything else is indexing.

Each function is written the way the reference applies it — one image, one step at a time — and the
device kernel (one fused gather) is checked against the composition.
"""
import numpy as np


def nearest_index(out_size: int, in_size: int) -> np.ndarray:
    """Source index of every output index (TF2 nearest, half-pixel centres, float32 arithmetic)."""
    scale = np.float32(in_size) / np.float32(out_size)
    o = np.arange(out_size, dtype=np.float32)
    idx = np.floor((o + np.float32(0.5)) * scale).astype(np.int64)
    return np.minimum(idx, in_size - 1)


def resize(image: np.ndarray, height: int, width: int) -> np.ndarray:
    """base_gan.py:45-53: tf.image.resize(image, [height, width], method=NEAREST_NEIGHBOR)."""
    return image[nearest_index(height, image.shape[0])][:, nearest_index(width, image.shape[1])]


def normalize(image: np.ndarray) -> np.ndarray:
    """base_gan.py:56-61: (image / 127.5) - 1 in float32."""
    return (image.astype(np.float32) / np.float32(127.5)) - np.float32(1.0)


def load(image_u8: np.ndarray, img_size=None) -> np.ndarray:
    """base_gan.py:26-44 after decoding: cast to float32, optional resize to img_size."""
    image = image_u8.astype(np.float32)
    return resize(image, img_size, img_size) if img_size else image


def split_img(image: np.ndarray, orient: str = 'left'):
    """pix2pix.py:34-55."""
    w = image.shape[1] // 2
    if orient == 'left':
        return image[:, :w, :], image[:, w:, :]
    return image[:, w:, :], image[:, :w, :]


def crop(image: np.ndarray, cy: int, cx: int, size: int) -> np.ndarray:
    """tf.image.random_crop with the offset already drawn (pix2pix.py:57-69, cycle_gan.py:38-45)."""
    return image[cy:cy + size, cx:cx + size, :]


def flip_left_right(image: np.ndarray) -> np.ndarray:
    return image[:, ::-1, :]


def pix2pix_process_train(pair_u8, orient, img_size, cy, cx, flip):
    """pix2pix.py:92-101 = split_img -> random_jitter (resize +30, crop, mirror) -> normalize."""
    a, b = split_img(load(pair_u8), orient)
    a, b = resize(a, img_size + 30, img_size + 30), resize(b, img_size + 30, img_size + 30)
    a, b = crop(a, cy, cx, img_size), crop(b, cy, cx, img_size)
    if flip:
        a, b = flip_left_right(a), flip_left_right(b)
    return normalize(a), normalize(b)


def pix2pix_process_pred(pair_u8, orient, img_size):
    """pix2pix.py:103-112."""
    a, b = split_img(load(pair_u8), orient)
    return normalize(resize(a, img_size, img_size)), normalize(resize(b, img_size, img_size))


def cyclegan_process_train(image_u8, img_size, cy, cx, flip):
    """cycle_gan.py:64-73: load(resize=True) -> resize +30 -> crop -> mirror -> normalize."""
    im = load(image_u8, img_size)
    im = crop(resize(im, img_size + 30, img_size + 30), cy, cx, img_size)
    if flip:
        im = flip_left_right(im)
    return normalize(im)


def cyclegan_process_pred(image_u8, img_size):
    """cycle_gan.py:75-85: load(resize=True) -> resize (identity) -> normalize."""
    return normalize(resize(load(image_u8, img_size), img_size, img_size))
