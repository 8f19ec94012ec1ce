"""GPU parity of the fused CycleGAN train step (reference cycle_gan.py:206-276) against the oracle,
which evaluates the reference's four separate tape.gradient calls; the device runs one sweep."""
import numpy as np
import pytest
import torch

from helpers import rel_err, load_model
from oracle import gan_oracle as O

pytestmark = pytest.mark.gpu
SEED = 321
CALLS = ['fake_y', 'cycled_x', 'fake_x', 'cycled_y', 'same_x', 'same_y']     # forward order, cycle_gan.py:220-228


def test_fp32_cyclegan_step_matches_oracle():
    from gan_b200 import CycleGAN
    cfg = dict(img_size=256, channels='3', learning_rate=2e-4, beta_1=0.5, beta_2=0.999, seed=SEED,
               precision='fp32', epochs=1, batch_size=1)
    cfg['lambda'] = 10
    m = CycleGAN(cfg)
    rng = np.random.default_rng(SEED + 1)
    specs = [O.generator_spec(3), O.generator_spec(3), O.discriminator_spec(3, False), O.discriminator_spec(3, False)]
    nets_np = [O.init_params(s, rng, "instancenorm") for s in specs]
    models = [m.generator_g, m.generator_f, m.discriminator_x, m.discriminator_y]
    for mod, arrs in zip(models, nets_np):
        load_model(mod, arrs)
    nets = [O.to_torch(a, torch.float64) for a in nets_np]
    irng = np.random.default_rng(SEED)
    b = 1
    x = O.synthetic_images(irng, b, 256, 256, 3); y = O.synthetic_images(irng, b, 256, 256, 3)
    c0 = m.ctx.call_counter()
    masks = {name: O.generator_keep_masks(SEED, c0 + i, 0, b, 256) for i, name in enumerate(CALLS)}
    losses = m.train_step(x, y, True)
    xt, yt = torch.tensor(x, dtype=torch.float64), torch.tensor(y, dtype=torch.float64)
    ref_losses, g1, g2, g3, g4, _ = O.cyclegan_losses_and_grads(nets[0], nets[1], nets[2], nets[3], xt, yt, 10.0, masks)
    nets32 = [O.to_torch(a, torch.float32) for a in nets_np]
    _, h1, h2, h3, h4, _ = O.cyclegan_losses_and_grads(nets32[0], nets32[1], nets32[2], nets32[3], torch.tensor(x),
                                                        torch.tensor(y), 10.0, masks)
    for a, r in zip(losses, ref_losses):
        assert abs(float(a) - float(r)) <= 1e-4 * max(1.0, abs(float(r))), (list(map(float, losses)), ref_losses)
    # gradients: within max(1e-4, 1.5x the float32 oracle's own deviation from float64) per tensor
    bad = []
    for mod, grads, grads32, tag in zip(models, (g1, g2, g3, g4), (h1, h2, h3, h4), "GFXY"):
        for v, g, g32 in zip(mod.trainable_variables, grads, grads32):
            g = g.numpy(); g32 = g32.numpy().astype(np.float64)
            den = np.abs(g).max()
            if den == 0.0:
                if np.abs(v.grad()).max() >= 1e-10:
                    bad.append((tag, v.name, "oracle exactly zero"))
                continue
            e, e32 = np.abs(v.grad() - g).max() / den, np.abs(g32 - g).max() / den
            if e > max(1e-4, 1.5 * e32):
                bad.append((tag, v.name, f"dev={e:.2e}", f"fp32-oracle={e32:.2e}"))
    assert not bad, bad
    assert all(o.iterations == 1 for o in (m.generator_g_optimizer, m.generator_f_optimizer,
                                           m.discriminator_x_optimizer, m.discriminator_y_optimizer))
    m.ctx.close()
