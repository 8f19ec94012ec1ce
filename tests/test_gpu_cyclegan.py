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
    opts = [O.KerasAdam(p) for p in nets]
    irng = np.random.default_rng(SEED)
    b = 1
    x = O.synthetic_images(irng, b, 256, 256, 3); y = O.synthetic_images(irng, b, 256, 256, 3)
    c0 = m.ctx.call_counter()
    masks = {name: O.generator_keep_masks(SEED, c0 + i, 0, b, 256) for i, name in enumerate(CALLS)}
    losses = m.train_step(x, y, True)
    ref_losses, ref_grads = O.cyclegan_train_step(nets, opts, torch.tensor(x, dtype=torch.float64),
                                                  torch.tensor(y, dtype=torch.float64), 10.0, True, masks)
    for a, r in zip(losses, ref_losses):
        assert abs(float(a) - r) <= 1e-4 * max(1.0, abs(r)), (list(map(float, losses)), ref_losses)
    for mod, grads, params, tag in zip(models, ref_grads, nets, "GFXY"):
        for v, g, p in zip(mod.trainable_variables, grads, params):
            g = g.numpy()
            if np.abs(g).max() == 0.0:
                assert np.abs(v.grad()).max() < 1e-10, (tag, v.name)
                continue
            assert rel_err(v.grad(), g) < 1e-4, (tag, v.name, "grad")
            assert rel_err(v.numpy(), p.detach().numpy()) < 1e-4, (tag, v.name, "weight")
    m.ctx.close()
