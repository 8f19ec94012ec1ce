"""bf16 / tcgen05 parity at the configurations BASELINE.json names and bench.py measures:

  * configs[1] at N=1: Pix2Pix 256x256 RGB, batch 64 — against tests/golden/pix2pix_b64_cal.npz (the float64
    oracle needs minutes per step at this size, so it ran once in the build container:
    tests/golden/make_golden_b64.py), with the float32 oracle's own free-running deviation beside it as
    the calibrator of the "after N steps" tolerance;
  * configs[3]: Pix2Pix 512x512 (1 channel), batch 4, live oracle;
  * configs[2]: CycleGAN 256x256 RGB, batch 4: generator outputs, seven losses, gradient cosines.

Metrics: max-rel = max|dev-ref| / max|ref| per tensor; L2 = ||dev-ref|| / ||ref||.
Reference sites: pix2pix.py:190-218, cycle_gan.py:206-276, base_gan.py:124-225.
"""
import os

import numpy as np
import pytest
import torch

from helpers import rel_err, make_pix2pix, load_model
from oracle import gan_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CALLS = ['fake_y', 'cycled_x', 'fake_x', 'cycled_y', 'same_x', 'same_y']


def _l2(a, b):
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))


def _cos(a, b):
    a = np.asarray(a, dtype=np.float64).ravel(); b = np.asarray(b, dtype=np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))


def _pix2pix(precision, channels, size, seed):
    from gan_b200 import Pix2Pix
    cfg = dict(img_size=size, channels=str(channels), learning_rate=2e-4, beta_1=0.5, beta_2=0.999,
               generator_loss='l1', seed=seed, precision=precision, epochs=1, batch_size=1)
    cfg['lambda'] = 100
    m = Pix2Pix(cfg)
    g_np, d_np = make_pix2pix(seed + 1, channels, None)
    load_model(m.generator, g_np); load_model(m.discriminator, d_np)
    return m, g_np, d_np


def test_bf16_batch64_benchmarked_config_against_golden():
    """BASELINE configs[1], N=1 (what bench.py times).  Protocol of make_golden_b64.py: out0 = G(x) [call 0],
    N=10 train steps [calls 1..10] through the CUDA-graph path bench.py uses, outN = G(x) [call 11]."""
    gold = np.load(os.path.join(HERE, "golden", "pix2pix_b64_cal.npz"))
    seed, B, size, ch, N = (int(v) for v in gold["protocol"])
    samples = [int(s) for s in gold["samples"]]
    m, g_np, d_np = _pix2pix("bf16", ch, size, seed)
    m.ctx.set_graphs(True)
    irng = np.random.default_rng(seed)
    x = O.synthetic_images(irng, B, size, size, ch); y = O.synthetic_images(irng, B, size, size, ch)
    assert m.ctx.call_counter() == 0
    # ---- step 0: identical weights ------------------------------------------------------------
    out0 = m.generator(x)
    ref0 = gold["out0_64"].astype(np.float64)
    mx0 = float(np.abs(out0[samples] - ref0).max() / gold["out0_norms"][0])
    l20 = float(np.linalg.norm(out0[samples] - ref0) / gold["out0_norms"][2])
    print(f"B=64 bf16 step 0: generator output max-rel={mx0:.3e} L2={l20:.3e}   "
          f"(float32 oracle vs float64 oracle: max-rel={gold['out0_f32_vs_f64'][0]:.3e} L2={gold['out0_f32_vs_f64'][1]:.3e})")
    assert mx0 <= 1e-2, mx0
    assert l20 <= 5e-3, l20
    # ---- N train steps ------------------------------------------------------------------------
    dev_losses = []
    for s in range(N):
        losses = [float(v) for v in m.train_step(x, y, True)]
        dev_losses.append(losses)
        if s == 0:
            cos = {}
            for tag, model in (("gg", m.generator), ("dg", m.discriminator)):
                off = 0
                for v, n in zip(model.trainable_variables, gold[tag + "_len"]):
                    idx = gold[tag + "_idx"][off:off + n]; ref = gold[tag + "_val"][off:off + n]; off += n
                    if v.name.endswith(".kernel") and np.abs(ref).max() > 0:
                        cos[f"{tag[0]}.{v.name}"] = _cos(v.grad().reshape(-1)[idx], ref)
            print("B=64 bf16 step-1 gradient cosines (4096-entry subsample per kernel):",
                  {k: round(c, 4) for k, c in cos.items()})
            assert min(cos.values()) > 0.95, cos
    dev_losses = np.array(dev_losses)
    l64, l32 = gold["losses64"], gold["losses32"]
    dev_dev = np.abs(dev_losses - l64) / np.maximum(1.0, np.abs(l64))
    f32_dev = np.abs(l32 - l64) / np.maximum(1.0, np.abs(l64))
    print("B=64 bf16 loss deviation from the float64 oracle per step (max over the 4 losses):",
          [f"{v:.2e}" for v in dev_dev.max(axis=1)])
    print("float32 ORACLE loss deviation from the float64 oracle per step                  :",
          [f"{v:.2e}" for v in f32_dev.max(axis=1)])
    assert dev_dev.max() <= 1e-2, dev_dev
    assert all(o.iterations == N for o in (m.generator_optimizer, m.discriminator_optimizer))
    # ---- after N steps -------------------------------------------------------------------------
    outN = m.generator(x)
    refN = gold["outN_64"].astype(np.float64)
    mx_free = float(np.abs(outN[samples] - refN).max() / gold["outN_norms"][0])
    l2_free = float(np.linalg.norm(outN[samples] - refN) / gold["outN_norms"][2])
    # oracle re-evaluated at the device's weights (same masks): isolates the forward arithmetic from the
    # +-lr sign flips of Keras-Adam's first updates
    masks = O.generator_keep_masks(seed, 1 + N, 0, B, size)
    with torch.no_grad():
        gp_dev = O.to_torch(m.generator.get_weights(), torch.float32, requires_grad=False)
        ref_sync = O.generator_forward(gp_dev, torch.tensor(x), "batchnorm", masks).numpy().astype(np.float64)
    mx_sync, l2_sync = rel_err(outN, ref_sync), _l2(outN, ref_sync)
    cal_mx, cal_l2 = (float(v) for v in gold["outN_f32_vs_f64"])
    print(f"B=64 bf16 after N={N} steps: free-running vs float64 oracle max-rel={mx_free:.3e} L2={l2_free:.3e}; "
          f"float32 ORACLE free-running vs float64 oracle max-rel={cal_mx:.3e} L2={cal_l2:.3e}; "
          f"device vs oracle at the device's weights max-rel={mx_sync:.3e} L2={l2_sync:.3e}")
    # the stated tolerance holds for the arithmetic (identical weights) ...
    assert mx_sync <= 1e-2 and l2_sync <= 5e-3, (mx_sync, l2_sync)
    # ... and the free-running trajectory is bounded by what torch float32 itself can hold (calibrated bound)
    assert l2_free <= max(1e-2, 3.0 * cal_l2), (l2_free, cal_l2)
    m.ctx.close()


@pytest.mark.parametrize("batch", [2, 8])
def test_bf16_step0_output_within_stated_tolerance(batch):
    """Step-0 generator output at B in {2, 8}: max-rel <= 1e-2 and L2 <= 5e-3 (B = 64 is covered above)."""
    m, g_np, _ = _pix2pix("bf16", 3, 256, 123)
    irng = np.random.default_rng(123)
    x = O.synthetic_images(irng, batch, 256, 256, 3)
    masks = O.generator_keep_masks(123, m.ctx.call_counter(), 0, batch, 256)
    out = m.generator(x)
    with torch.no_grad():
        ref = O.generator_forward(O.to_torch(g_np, torch.float64, False), torch.tensor(x, dtype=torch.float64),
                                  "batchnorm", masks).numpy()
    mx, l2 = rel_err(out, ref), _l2(out, ref)
    print(f"bf16 step 0, B={batch}: max-rel={mx:.3e} L2={l2:.3e}")
    assert mx <= 1e-2 and l2 <= 5e-3, (mx, l2)
    m.ctx.close()


def test_bf16_pix2pix_512_batch4():
    """BASELINE configs[3] shard (512x512, 1 channel, 4 images per GPU) in the tcgen05 path: generator output,
    62x62 logits, four losses and gradient cosines against the live float64 oracle."""
    B = 4
    m, g_np, d_np = _pix2pix("bf16", 1, 512, 77)
    irng = np.random.default_rng(77)
    x = O.synthetic_images(irng, B, 512, 512, 1); y = O.synthetic_images(irng, B, 512, 512, 1)
    xt, yt = torch.tensor(x, dtype=torch.float64), torch.tensor(y, dtype=torch.float64)
    gp, dp = O.to_torch(g_np, torch.float64), O.to_torch(d_np, torch.float64)
    masks = O.generator_keep_masks(77, m.ctx.call_counter(), 0, B, 512)
    out = m.generator(x)
    with torch.no_grad():
        ref = O.generator_forward(gp, xt, "batchnorm", masks).numpy()
    mx, l2 = rel_err(out, ref), _l2(out, ref)
    print(f"bf16 512^2 B=4 step 0: generator output max-rel={mx:.3e} L2={l2:.3e}")
    assert mx <= 1e-2 and l2 <= 5e-3, (mx, l2)
    logits = m.discriminator([x, y])
    with torch.no_grad():
        ref_l = O.discriminator_forward(dp, xt, yt).numpy()
    assert logits.shape == (B, 62, 62, 1) and rel_err(logits, ref_l) <= 1e-2, rel_err(logits, ref_l)
    masks = O.generator_keep_masks(77, m.ctx.call_counter(), 0, B, 512)
    losses = m.train_step(x, y, True)
    ref_losses, gg, dg, _ = O.pix2pix_losses_and_grads(gp, dp, xt, yt, 100.0, masks)
    for a, r in zip(losses, ref_losses):
        assert abs(float(a) - float(r)) <= 1e-2 * max(1.0, abs(float(r))), (list(map(float, losses)), ref_losses)
    cos = {}
    for tag, model, grads in (("g", m.generator, gg), ("d", m.discriminator, dg)):
        for v, g in zip(model.trainable_variables, grads):
            if v.name.endswith(".kernel") and float(g.abs().max()) > 0:
                cos[f"{tag}.{v.name}"] = _cos(v.grad(), g.numpy())
    print("bf16 512^2 gradient cosines:", {k: round(c, 4) for k, c in cos.items()})
    assert min(cos.values()) > 0.95, cos
    m.ctx.close()


def test_bf16_cyclegan_batch4_outputs_losses_and_gradients():
    """BASELINE configs[2] shard (CycleGAN 256x256 RGB, 4 pairs per GPU, InstanceNorm) in the tcgen05 path:
    all six generator outputs, seven losses, and the gradient direction of every conv kernel of the four nets
    (single backward sweep on the device == the reference's four tape.gradient calls, cycle_gan.py:252-260)."""
    from gan_b200 import CycleGAN
    seed, B = 321, 4
    cfg = dict(img_size=256, channels='3', learning_rate=2e-4, beta_1=0.5, beta_2=0.999, seed=seed, precision='bf16')
    cfg['lambda'] = 10
    m = CycleGAN(cfg)
    rng = np.random.default_rng(seed + 1)
    specs = [O.generator_spec(3), O.generator_spec(3), O.discriminator_spec(3, False), O.discriminator_spec(3, False)]
    nets_np = [O.init_params(s, rng, "instancenorm") for s in specs]
    models = [m.generator_g, m.generator_f, m.discriminator_x, m.discriminator_y]
    for mod, arrs in zip(models, nets_np):
        load_model(mod, arrs)
    nets = [O.to_torch(a, torch.float64) for a in nets_np]
    irng = np.random.default_rng(seed)
    x = O.synthetic_images(irng, B, 256, 256, 3); y = O.synthetic_images(irng, B, 256, 256, 3)
    c0 = m.ctx.call_counter()
    masks = {n: O.generator_keep_masks(seed, c0 + i, 0, B, 256) for i, n in enumerate(CALLS)}
    losses = m.train_step(x, y, True)
    ref, g1, g2, g3, g4, outs = O.cyclegan_losses_and_grads(nets[0], nets[1], nets[2], nets[3],
                                                             torch.tensor(x, dtype=torch.float64),
                                                             torch.tensor(y, dtype=torch.float64), 10.0, masks)
    for a, r in zip(losses, ref):
        assert abs(float(a) - float(r)) <= 1e-2 * max(1.0, abs(float(r))), (list(map(float, losses)), [float(v) for v in ref])
    # generator outputs saved by the step: slots follow the forward order per net (engine.cu cyclegan_step)
    slot_of = {"fake_y": (m.generator_g, 0), "cycled_x": (m.generator_f, 0), "fake_x": (m.generator_f, 1),
               "cycled_y": (m.generator_g, 1), "same_x": (m.generator_f, 2), "same_y": (m.generator_g, 2)}
    worst = {}
    for name, (model, slot) in slot_of.items():
        dev = model.debug_tensor("out", slot=slot)
        r = outs[name].detach().numpy().reshape(-1)
        worst[name] = (rel_err(dev, r), _l2(dev, r))
    print("bf16 CycleGAN B=4 generator outputs (max-rel, L2):", {k: (f"{a:.2e}", f"{b:.2e}") for k, (a, b) in worst.items()})
    # first-generation outputs see one generator; cycled outputs stack two (32 bf16 layers)
    for name, (mx, l2) in worst.items():
        lim = 2.0 if name.startswith("cycled") else 1.0
        assert mx <= lim * 1e-2 and l2 <= lim * 5e-3, (name, mx, l2)
    cos = {}
    for tag, model, grads in zip("GFXY", models, (g1, g2, g3, g4)):
        for v, g in zip(model.trainable_variables, grads):
            if v.name.endswith(".kernel") and float(g.abs().max()) > 0:
                cos[f"{tag}.{v.name}"] = _cos(v.grad(), g.numpy())
    print("bf16 CycleGAN gradient cosines:", {k: round(c, 4) for k, c in cos.items()})
    assert min(cos.values()) > 0.95, cos
    m.ctx.close()
