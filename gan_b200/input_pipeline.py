"""Host side of the on-device input pipeline.

The reference maps every image through ``split_img`` / ``resize`` / ``random_crop`` /
``flip_left_right`` / ``normalize`` on the CPU inside tf.data (base_gan.py:26-61, pix2pix.py:34-112,
cycle_gan.py:38-85).  Here the host only decodes files to uint8 (PIL), draws the random crop offset
and flip, and describes each image by a ``gan_image_xform``; every pixel operation is ONE gather
kernel on the device (``gan_preprocess_images`` / ``gan_ctx_prefetch_images``), and the host->device
traffic is the uint8 bytes.  There is no host implementation of the pixel arithmetic.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _ffi

JITTER = 30        # resize to img_size + 30 before the random crop (pix2pix.py:77-78, cycle_gan.py:53)


def load(image_file: str, channels: int) -> np.ndarray:
    """Reference base_gan.py:26-44 without the float cast: decode PNG/JPEG to (H, W, C) uint8."""
    from PIL import Image
    with Image.open(image_file) as im:
        im = im.convert({1: "L", 3: "RGB", 4: "RGBA"}[int(channels)])
        a = np.asarray(im, dtype=np.uint8)
    return a[:, :, None] if a.ndim == 2 else a


def draw_jitter(rng: np.random.Generator, img_size: int):
    """The random draws of ``random_jitter`` (pix2pix.py:66-90): crop offset uniform in [0, 30] per
    axis (tf.image.random_crop), mirror with probability 1/2."""
    cy, cx = (int(v) for v in rng.integers(0, JITTER + 1, size=2))
    return cy, cx, bool(rng.random() > 0.5)


def xform(src_h, src_w, col0, cols, img_size, train, crop=(0, 0), flip=False, pre=0) -> _ffi.ImageXform:
    """One image's transform.  train: resize to img_size+30, crop, flip (process_images_train);
    otherwise resize to img_size only (process_images_pred)."""
    if train:
        return _ffi.ImageXform(src_h, src_w, col0, cols, pre, img_size + JITTER, int(crop[0]), int(crop[1]), int(bool(flip)))
    return _ffi.ImageXform(src_h, src_w, col0, cols, pre, 0, 0, 0, 0)


def pack_images(images):
    """List of (H, W, C) uint8 arrays (sizes may differ) -> one contiguous byte buffer + stride."""
    stride = max(int(im.size) for im in images)
    buf = np.zeros((len(images), stride), dtype=np.uint8)
    for i, im in enumerate(images):
        if im.dtype != np.uint8 or im.ndim != 3:
            raise ValueError("images must be decoded (H, W, C) uint8 arrays")
        buf[i, :im.size] = np.ascontiguousarray(im).reshape(-1)
    return buf, stride


def _addr(t):
    """Address of a uint8 buffer: numpy array, or torch tensor (pinned host or CUDA)."""
    return C.c_void_p(t.data_ptr() if hasattr(t, "data_ptr") else t.ctypes.data)


def _xf_array(xfs):
    arr = (_ffi.ImageXform * len(xfs))(*xfs)
    return arr


def preprocess(ctx, packed, xfs, channels: int, img_size: int, out=None):
    """Synchronous form: ``packed`` = (uint8 buffer, stride) from ``pack_images``.  Returns
    (B, S, S, C) float32 in [-1, 1] (numpy, or fills ``out`` which may be a CUDA torch tensor — then the
    call only enqueues the work on the context stream)."""
    buf, stride = packed
    b = len(xfs)
    if out is None:
        out = np.empty((b, img_size, img_size, channels), dtype=np.float32)
    _ffi.check(_ffi.lib().gan_preprocess_images(ctx.handle, _addr(buf), stride, b, channels, img_size,
                                                C.cast(_xf_array(xfs), C.c_void_p), _ffi.ptr_of(out)))
    return out


class DeviceBatch:
    """Device-resident float32 image batch produced by ``prefetch`` (a borrowed pointer into the
    context's prefetch buffers: valid until the train step that consumes it has been enqueued)."""

    def __init__(self, ptr: int, shape):
        self.ptr, self.shape = ptr, tuple(shape)
        self.dtype = np.float32

    def data_ptr(self):
        return self.ptr


def prefetch(ctx, buf_a, stride_a, xfs_a, buf_b, stride_b, xfs_b, channels: int, img_size: int):
    """Asynchronous form (tf.data ``prefetch`` role): uint8 copy + gather on the copy stream.  Returns
    two ``DeviceBatch`` objects to hand to the next ``train_step``.  ``buf_*`` must stay alive (and
    should be pinned) until that step has been enqueued."""
    b = len(xfs_a)
    pa, pb = C.c_void_p(), C.c_void_p()
    _ffi.check(_ffi.lib().gan_ctx_prefetch_images(ctx.handle, _addr(buf_a), stride_a, C.cast(_xf_array(xfs_a), C.c_void_p),
                                                  _addr(buf_b), stride_b, C.cast(_xf_array(xfs_b), C.c_void_p), b, channels,
                                                  img_size, C.byref(pa), C.byref(pb)))
    shape = (b, img_size, img_size, channels)
    return DeviceBatch(pa.value, shape), DeviceBatch(pb.value, shape)


def list_images(path: str):
    """Reference pix2pix.py:130 / cycle_gan.py:96: files whose name contains 'png' or 'jpg'."""
    return [i for i in os.listdir(path) if 'png' in i or 'jpg' in i]
