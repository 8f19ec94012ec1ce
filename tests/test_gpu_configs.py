"""Parity cases for the remaining BASELINE.json configs: Pix2Pix 512x512 (2x2 bottleneck, 62x62
logits), generator-only predict at large batch, CycleGAN in the bf16/tcgen05 path."""
import numpy as np
import pytest
import torch

from helpers import rel_err, load_model
from oracle import gan_oracle as O

pytestmark = pytest.mark.gpu
SEED = 77


def _l2(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))


def test_pix2pix_512_step_matches_oracle():
    """BASELINE config 4 (thermal->visible default: 512x512, 1 channel).  fp32 path, batch 1."""
    from gan_b200 import Pix2Pix
    cfg = dict(img_size=512, channels='1', learning_rate=2e-4, beta_1=0.5, beta_2=0.999, generator_loss='l1',
               seed=SEED, precision='fp32')
    cfg['lambda'] = 100
    m = Pix2Pix(cfg)
    rng = np.random.default_rng(SEED + 1)
    g_np = O.init_params(O.generator_spec(1), rng, "batchnorm")
    d_np = O.init_params(O.discriminator_spec(1, True), rng, "batchnorm")
    load_model(m.generator, g_np); load_model(m.discriminator, d_np)
    irng = np.random.default_rng(SEED)
    x = O.synthetic_images(irng, 1, 512, 512, 1); y = O.synthetic_images(irng, 1, 512, 512, 1)
    masks = O.generator_keep_masks(SEED, m.ctx.call_counter(), 0, 1, 512)
    logits = m.discriminator([x, y])
    assert logits.shape == (1, 62, 62, 1)                       # 64 -> 63 -> 62
    gp, dp = O.to_torch(g_np, torch.float64), O.to_torch(d_np, torch.float64)
    ref_logits = O.discriminator_forward(dp, torch.tensor(x, dtype=torch.float64), torch.tensor(y, dtype=torch.float64))
    assert rel_err(logits, ref_logits.detach().numpy()) < 1e-4
    losses = m.train_step(x, y, True)
    ref, gg, dg, gen_out = O.pix2pix_losses_and_grads(gp, dp, torch.tensor(x, dtype=torch.float64),
                                                      torch.tensor(y, dtype=torch.float64), 100.0, masks)
    for a, r in zip(losses, ref):
        assert abs(float(a) - float(r)) <= 1e-4 * max(1.0, abs(float(r))), (list(map(float, losses)), ref)
    gp32, dp32 = O.to_torch(g_np, torch.float32), O.to_torch(d_np, torch.float32)
    _, gg32, _, _ = O.pix2pix_losses_and_grads(gp32, dp32, torch.tensor(x), torch.tensor(y), 100.0, masks)
    for v, g, g32 in zip(m.generator.trainable_variables, gg, gg32):
        g = g.numpy(); den = np.abs(g).max()
        if den == 0:
            continue
        e, e32 = np.abs(v.grad() - g).max() / den, np.abs(g32.numpy().astype(np.float64) - g).max() / den
        assert e <= max(1e-4, 1.5 * e32), (v.name, e, e32)
    m.ctx.close()


def test_generator_predict_large_batch():
    """BASELINE config 5: generator forward with training=True semantics (pix2pix.py:228) at batch 32
    against the oracle (bf16 path), and at batch 256 as a run: bounded by tanh, finite, and
    reproducible when the dropout counter is rewound."""
    from gan_b200 import Pix2Pix
    cfg = dict(img_size=256, channels='3', seed=SEED, precision='bf16')
    m = Pix2Pix(cfg)
    rng = np.random.default_rng(SEED + 1)
    g_np = O.init_params(O.generator_spec(3), rng, "batchnorm")
    load_model(m.generator, g_np)
    irng = np.random.default_rng(SEED)
    x = O.synthetic_images(irng, 32, 256, 256, 3)
    masks = O.generator_keep_masks(SEED, m.ctx.call_counter(), 0, 32, 256)
    out = m.generator(x, training=True)
    ref = O.generator_forward(O.to_torch(g_np, torch.float32), torch.tensor(x), "batchnorm", masks).detach().numpy()
    assert _l2(out, ref.astype(np.float64)) < 1e-2
    xb = O.synthetic_images(irng, 256, 256, 256, 3)
    m.ctx.set_rng(SEED, 100)
    o1 = m.generator(xb)
    m.ctx.set_rng(SEED, 100)
    o2 = m.generator(xb)
    assert o1.shape == (256, 256, 256, 3) and np.isfinite(o1).all() and np.abs(o1).max() <= 1.0
    assert np.array_equal(o1, o2)                                # forward path is deterministic
    m.ctx.set_rng(SEED, 101)
    assert not np.array_equal(o1, m.generator(xb))              # a different call counter draws different dropout masks
    m.ctx.close()


def test_cyclegan_bf16_step_tracks_oracle():
    """BASELINE config 3 in the tcgen05 path: seven losses within 1e-2 of the float64 oracle."""
    from gan_b200 import CycleGAN
    cfg = dict(img_size=256, channels='3', learning_rate=2e-4, beta_1=0.5, beta_2=0.999, seed=SEED, precision='bf16')
    cfg['lambda'] = 10
    m = CycleGAN(cfg)
    rng = np.random.default_rng(SEED + 1)
    specs = [O.generator_spec(3), O.generator_spec(3), O.discriminator_spec(3, False), O.discriminator_spec(3, False)]
    nets_np = [O.init_params(s, rng, "instancenorm") for s in specs]
    for mod, arrs in zip([m.generator_g, m.generator_f, m.discriminator_x, m.discriminator_y], nets_np):
        load_model(mod, arrs)
    nets = [O.to_torch(a, torch.float64) for a in nets_np]
    irng = np.random.default_rng(SEED)
    b = 2
    x = O.synthetic_images(irng, b, 256, 256, 3); y = O.synthetic_images(irng, b, 256, 256, 3)
    calls = ['fake_y', 'cycled_x', 'fake_x', 'cycled_y', 'same_x', 'same_y']
    c0 = m.ctx.call_counter()
    masks = {n: O.generator_keep_masks(SEED, c0 + i, 0, b, 256) for i, n in enumerate(calls)}
    losses = m.train_step(x, y, True)
    ref, _, _, _, _, _ = O.cyclegan_losses_and_grads(nets[0], nets[1], nets[2], nets[3], torch.tensor(x, dtype=torch.float64),
                                                     torch.tensor(y, dtype=torch.float64), 10.0, masks, want_grads=False)
    for a, r in zip(losses, ref):
        assert abs(float(a) - float(r)) <= 1e-2 * max(1.0, abs(float(r))), (list(map(float, losses)), [float(v) for v in ref])
    # validation step: no update
    w = m.generator_g.get_flat_params()
    m.train_step(x, y, False)
    assert np.array_equal(w, m.generator_g.get_flat_params())
    m.ctx.close()
