"""Edge cases of the train-step path through the C-ABI: odd and ragged batches, error behaviour
(bad arguments come back as error codes with a message, nothing throws across the boundary and the
context stays usable), run-to-run determinism of the forward pass."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import make_pix2pix, load_model
from oracle import gan_oracle as O

pytestmark = pytest.mark.gpu

SEED = 123


def _build(precision, channels=3):
    from gan_b200 import Pix2Pix
    cfg = dict(img_size=256, channels=str(channels), learning_rate=2e-4, beta_1=0.5, beta_2=0.999,
               generator_loss='l1', seed=SEED, precision=precision, epochs=1, batch_size=1)
    cfg['lambda'] = 100
    m = Pix2Pix(cfg)
    g_np, d_np = make_pix2pix(SEED + 1, channels, None)
    load_model(m.generator, g_np)
    load_model(m.discriminator, d_np)
    return m, g_np, d_np


def _inputs(b, c, seed=SEED):
    rng = np.random.default_rng(seed)
    return O.synthetic_images(rng, b, 256, 256, c), O.synthetic_images(rng, b, 256, 256, c)


def test_bf16_odd_batch_losses_track_oracle():
    """Batch 3 on the tcgen05 path: pixel counts are not powers of two, M tiles straddle samples, the
    small-layer BatchNorm kernel and the split-K slabs see ragged sizes.  Losses within 1e-2 of the
    float64 oracle (BASELINE tolerance of the bf16 path)."""
    m, g_np, d_np = _build("bf16")
    x, y = _inputs(3, 3, seed=5)
    masks = O.generator_keep_masks(SEED, m.ctx.call_counter(), 0, 3, 256)
    losses = m.train_step(x, y, False)
    gp, dp = O.to_torch(g_np, torch.float64), O.to_torch(d_np, torch.float64)
    ref, _, _ = O.pix2pix_train_step(gp, dp, O.KerasAdam(gp), O.KerasAdam(dp), torch.tensor(x, dtype=torch.float64),
                                     torch.tensor(y, dtype=torch.float64), 100.0, False, masks)
    for a, r in zip(losses, ref):
        assert abs(float(a) - r) <= 1e-2 * max(1.0, abs(r)), (list(map(float, losses)), ref)
    m.ctx.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_validation_step_is_bitwise_repeatable_and_batch_independent_of_history(precision):
    """training=False steps do not touch any state except the dropout call counter: with the counter
    rewound the same step returns bit-identical losses, also after steps of other batch sizes ran in
    between (buffers are re-used across batch sizes)."""
    m, _, _ = _build(precision)
    x, y = _inputs(2, 3, seed=9)
    c0 = m.ctx.call_counter()
    a = [float(v) for v in m.train_step(x, y, False)]
    m.train_step(x[:1], y[:1], False)                     # ragged tail
    x3, y3 = _inputs(3, 3, seed=10)
    m.train_step(x3, y3, False)
    m.ctx.set_rng(SEED, c0)
    b = [float(v) for v in m.train_step(x, y, False)]
    assert a == b
    m.ctx.close()


def test_bad_arguments_return_error_codes_and_leave_the_context_usable():
    from gan_b200 import _ffi
    m, _, _ = _build("bf16")
    lib = _ffi.lib()
    x, y = _inputs(1, 3)
    losses = np.zeros(4, dtype=np.float32)
    g_opt = m.generator_optimizer.bind(m.generator)
    d_opt = m.discriminator_optimizer.bind(m.discriminator)
    args = (m.generator.handle, m.discriminator.handle, g_opt, d_opt)
    # batch 0
    rc = lib.gan_pix2pix_train_step(*args, _ffi.ptr_of(x), _ffi.ptr_of(y), 0, C.c_float(100.0), 1, _ffi.ptr_of(losses))
    assert rc == -1 and b"batch" in lib.gan_last_error()
    # null image pointer
    rc = lib.gan_pix2pix_train_step(*args, None, _ffi.ptr_of(y), 1, C.c_float(100.0), 1, _ffi.ptr_of(losses))
    assert rc == -1
    # generator and discriminator swapped
    rc = lib.gan_pix2pix_train_step(m.discriminator.handle, m.generator.handle, g_opt, d_opt, _ffi.ptr_of(x), _ffi.ptr_of(y),
                                    1, C.c_float(100.0), 1, _ffi.ptr_of(losses))
    assert rc == -1
    # optimizers bound to the wrong nets
    rc = lib.gan_pix2pix_train_step(m.generator.handle, m.discriminator.handle, d_opt, g_opt, _ffi.ptr_of(x), _ffi.ptr_of(y),
                                    1, C.c_float(100.0), 1, _ffi.ptr_of(losses))
    assert rc == -1
    # unsupported geometry at construction: the U-Net has 8 stride-2 levels (base_gan.py:180-189)
    bad = C.c_void_p()
    assert lib.gan_generator_create(m.ctx.handle, 1, 200, 200, 3, C.byref(bad)) == -1
    assert lib.gan_generator_create(m.ctx.handle, 1, 256, 256, 7, C.byref(bad)) == -1
    assert lib.gan_net_get_tensor(m.generator.handle, 10_000, None) == -1
    with pytest.raises(ValueError):
        m.train_step(x, np.zeros((1, 256, 256, 1), np.float32))
    with pytest.raises(_ffi.GanError):
        _ffi.check(-1)
    # the context still trains after the rejected calls
    out = m.train_step(x, y, True)
    assert all(np.isfinite(float(v)) for v in out)
    m.ctx.close()


def test_graph_replay_survives_buffer_growth_from_a_larger_batch():
    """Captured CUDA graphs hold raw device pointers.  A later, larger batch re-allocates the activation
    buffers; the small-batch graph must then be re-captured, not replayed with stale addresses (the ragged
    last batch of an epoch followed by full batches is exactly this sequence)."""
    m, _, _ = _build("bf16")
    m.ctx.set_graphs(True)
    x2, y2 = _inputs(2, 3, seed=31)
    x4, y4 = _inputs(4, 3, seed=32)
    c0 = m.ctx.call_counter()
    first = None
    for _ in range(3):                                    # eager, capture, replay at batch 2
        m.ctx.set_rng(SEED, c0)                           # (set_rng drops the graphs: each pass re-captures) 
        first = [float(v) for v in m.train_step(x2, y2, False)]
    a = [float(v) for v in m.train_step(x2, y2, False)]   # eager
    b = [float(v) for v in m.train_step(x2, y2, False)]   # capture + launch
    for _ in range(3):
        m.train_step(x4, y4, False)                       # grows every activation buffer
    c_before = m.ctx.call_counter()
    got = [float(v) for v in m.train_step(x2, y2, False)]         # must not replay the stale graph
    m.ctx.set_graphs(False)
    m.ctx.set_rng(SEED, c_before)
    want = [float(v) for v in m.train_step(x2, y2, False)]        # eager reference at the same dropout counter
    assert got == want and all(np.isfinite(v) for v in a + b + first)
    m.ctx.close()


def test_handle_ownership_and_double_destroy():
    """The context owns its nets and optimizers (include/gan_b200.h): destroying a net takes its optimizer
    with it, a second destroy of any handle is an error code, closing the context frees the rest."""
    from gan_b200 import _ffi
    m, _, _ = _build("bf16")
    lib = _ffi.lib()
    x, y = _inputs(1, 3)
    m.train_step(x, y, True)                              # binds both optimizers
    g, g_opt = m.generator.handle, m.generator_optimizer._h
    d_opt = m.discriminator_optimizer._h
    assert lib.gan_adam_destroy(d_opt) == 0 and lib.gan_adam_destroy(d_opt) == -1
    assert lib.gan_net_destroy(g) == 0                    # takes g_opt with it
    assert lib.gan_adam_destroy(g_opt) == -1 and lib.gan_net_destroy(g) == -1
    h = m.ctx.handle
    m.ctx.close()                                         # frees the discriminator
    assert lib.gan_ctx_destroy(h) == -1


def test_ctx_setters_invalidate_captured_graphs():
    """gan_ctx_set_dropout / set_sample_offset change state a captured step graph has baked in (kernel template,
    kernel argument): the setters must drop the graphs so the next step honours the new value."""
    m, _, _ = _build("bf16")
    m.ctx.set_graphs(True)
    x, y = _inputs(2, 3, seed=41)
    for _ in range(4):                                    # eager, capture, replay, replay
        m.train_step(x, y, False)
    c = m.ctx.call_counter()
    with_drop = [float(v) for v in m.train_step(x, y, False)]             # replayed graph, dropout on
    m.ctx.set_dropout(False)                                              # after capture: must not be ignored
    runs = [[float(v) for v in m.train_step(x, y, False)] for _ in range(3)]   # eager, capture, replay
    ref = _build("bf16")[0]
    ref.ctx.set_dropout(False)
    want = [float(v) for v in ref.train_step(x, y, False)]                # eager, dropout off from the start
    assert all(r == want for r in runs) and want != with_drop
    # sample offset: a different global sample index draws different masks
    m.ctx.set_dropout(True)
    for _ in range(3):
        m.ctx.set_rng(SEED, c)
        a = [float(v) for v in m.train_step(x, y, False)]
    m.ctx.set_sample_offset(2)
    m.ctx.set_rng(SEED, c)
    b = [float(v) for v in m.train_step(x, y, False)]
    assert a == with_drop and b != a
    m.ctx.close(); ref.ctx.close()
