"""On-device input pipeline (gan_preprocess_images / gan_ctx_prefetch_images) against the per-image
CPU oracle (oracle/pipeline_oracle.py restating base_gan.py:26-61, pix2pix.py:34-112,
cycle_gan.py:38-85).  Byte/index work: the bar is bit-exact."""
import ctypes as C

import numpy as np
import pytest

from oracle import pipeline_oracle as P

pytestmark = pytest.mark.gpu

SEED = 123


def _p2p(channels=3, orient='left', precision='bf16'):
    from gan_b200 import Pix2Pix
    cfg = dict(img_size=256, channels=str(channels), learning_rate=2e-4, beta_1=0.5, beta_2=0.999, generator_loss='l1',
               seed=SEED, precision=precision, epochs=1, batch_size=1, input_img_orient=orient)
    cfg['lambda'] = 100
    return Pix2Pix(cfg)


def _pairs(rng, sizes, c):
    return [rng.integers(0, 256, size=(h, w, c), dtype=np.uint8) for h, w in sizes]


@pytest.mark.parametrize("channels,orient", [(3, 'left'), (3, 'right'), (1, 'left')])
def test_pix2pix_train_and_pred_pipeline_bit_exact(channels, orient):
    """Ragged batch: every pair has its own size (odd widths make the two halves differ by a column),
    the draws cover both corners of the crop range and both mirror states."""
    m = _p2p(channels, orient)
    rng = np.random.default_rng(7)
    pairs = _pairs(rng, [(256, 512), (300, 500), (257, 513), (90, 1023), (640, 1280)], channels)

    class Draws:                         # fixed draws in the order draw_jitter consumes them
        def __init__(self, seq): self.seq = list(seq); self.i = 0
        def integers(self, lo, hi, size): v = self.seq[self.i]; self.i += 1; return np.array(v[:2])
        def random(self): return 0.9 if self.seq[self.i - 1][2] else 0.1
    draws = [(0, 0, False), (30, 30, True), (17, 3, True), (30, 0, False), (5, 29, True)]
    a, b = m.process_images(pairs, True, Draws(draws))
    assert a.shape == (5, 256, 256, channels) and a.dtype == np.float32
    for n, (pair, (cy, cx, fl)) in enumerate(zip(pairs, draws)):
        ra, rb = P.pix2pix_process_train(pair, orient, 256, cy, cx, fl)
        assert np.array_equal(a[n], ra) and np.array_equal(b[n], rb), n
    a, b = m.process_images(pairs, False)
    for n, pair in enumerate(pairs):
        ra, rb = P.pix2pix_process_pred(pair, orient, 256)
        assert np.array_equal(a[n], ra) and np.array_equal(b[n], rb), n
    m.ctx.close()


def test_cyclegan_pipeline_bit_exact():
    from gan_b200 import CycleGAN
    cfg = dict(img_size=256, channels='3', learning_rate=2e-4, beta_1=0.5, beta_2=0.999, seed=SEED, precision='bf16',
               epochs=1, batch_size=1)
    cfg['lambda'] = 10
    m = CycleGAN(cfg)
    rng = np.random.default_rng(11)
    images = _pairs(rng, [(256, 256), (480, 640), (199, 301)], 3)
    jit = np.random.default_rng(5)
    out = m.process_images(images, True, jit)
    jit = np.random.default_rng(5)
    from gan_b200 import input_pipeline
    for n, im in enumerate(images):
        cy, cx, fl = input_pipeline.draw_jitter(jit, 256)
        assert np.array_equal(out[n], P.cyclegan_process_train(im, 256, cy, cx, fl)), n
    out = m.process_images(images, False)
    for n, im in enumerate(images):
        assert np.array_equal(out[n], P.cyclegan_process_pred(im, 256)), n
    m.ctx.close()


def test_prefetched_device_batches_feed_the_train_step():
    """gan_ctx_prefetch_images -> train_step(device batches) gives bit-identical losses to
    train_step(host float32 arrays of the synchronous pipeline) with the dropout counter rewound;
    covers eager and graph-replayed steps."""
    import torch
    from gan_b200 import input_pipeline
    m = _p2p(3, 'left')
    rng = np.random.default_rng(3)
    pairs = _pairs(rng, [(256, 512), (286, 572)], 3)
    x, y = m.process_images(pairs, True, np.random.default_rng(21))
    buf, stride = input_pipeline.pack_images(pairs)
    pinned = torch.from_numpy(buf).pin_memory()
    for it in range(3):                                   # eager, capture, replay
        c0 = m.ctx.call_counter()
        want = [float(v) for v in m.train_step(x, y, False)]
        m.ctx.set_rng(SEED, c0)
        dx, dy = m.prefetch_pairs((pinned, stride), pairs, True, np.random.default_rng(21))
        got = [float(v) for v in m.train_step(dx, dy, False)]
        assert got == want, (it, got, want)
    m.ctx.close()


def test_bad_transforms_are_rejected():
    from gan_b200 import _ffi
    m = _p2p(3)
    img = np.zeros((64, 128, 3), np.uint8)
    out = np.empty((1, 256, 256, 3), np.float32)
    lib = _ffi.lib()

    def call(xf):
        arr = (_ffi.ImageXform * 1)(xf)
        return lib.gan_preprocess_images(m.ctx.handle, C.c_void_p(img.ctypes.data), img.size, 1, 3, 256,
                                         C.cast(arr, C.c_void_p), _ffi.ptr_of(out))
    assert call(_ffi.ImageXform(64, 128, 0, 64, 0, 286, 0, 0, 0)) == 0
    assert call(_ffi.ImageXform(64, 128, 0, 64, 0, 286, 31, 0, 0)) == -1          # crop beyond the +30 border
    assert call(_ffi.ImageXform(64, 128, 100, 64, 0, 286, 0, 0, 0)) == -1         # window outside the image
    assert call(_ffi.ImageXform(64, 256, 0, 64, 0, 286, 0, 0, 0)) == -1           # image larger than its stride
    assert call(_ffi.ImageXform(64, 128, 0, 64, 0, 100, 0, 0, 0)) == -1           # resize target smaller than the crop
    assert b"" != lib.gan_last_error()
    m.ctx.close()
