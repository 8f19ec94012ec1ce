"""GPU parity of the fused Pix2Pix train step (reference pix2pix.py:190-218) against the CPU oracle:
identical host-generated weights, synthetic U[-1,1) images and Philox dropout masks."""
import numpy as np
import pytest
import torch

from helpers import rel_err, make_pix2pix, load_model
from oracle import gan_oracle as O

pytestmark = pytest.mark.gpu

SEED = 123
GRAD_FACTOR = 1.5     # device fp32 gradients: within 1.5x of torch-CPU fp32's own deviation from float64 (measured ratio 0.0-1.01)


def _build(precision, channels=3, size=256, lam=100):
    from gan_b200 import Pix2Pix
    cfg = dict(img_size=size, channels=str(channels), learning_rate=2e-4, beta_1=0.5, beta_2=0.999,
               generator_loss='l1', seed=SEED, precision=precision, epochs=1, batch_size=1)
    cfg['lambda'] = lam
    m = Pix2Pix(cfg)
    g_np, d_np = make_pix2pix(SEED + 1, channels, None)
    load_model(m.generator, g_np)
    load_model(m.discriminator, d_np)
    return m, g_np, d_np


def _oracle_state(g_np, d_np, dtype=torch.float64):
    gp, dp = O.to_torch(g_np, dtype), O.to_torch(d_np, dtype)
    return gp, dp, O.KerasAdam(gp), O.KerasAdam(dp)


def _inputs(b, size, c, seed=SEED):
    rng = np.random.default_rng(seed)
    return O.synthetic_images(rng, b, size, size, c), O.synthetic_images(rng, b, size, size, c)


def _check_tensors(names, dev, ref, tol, what, atol=0.0):
    """Per-tensor max|dev-ref| <= tol*max|ref| + atol; reports every offending tensor at once."""
    bad, worst = [], 0.0
    for n, a, r in zip(names, dev, ref):
        r = r.detach().numpy() if hasattr(r, "detach") else np.asarray(r)
        if np.abs(r).max() == 0.0:
            if np.abs(a).max() >= 1e-10:
                bad.append((n, "oracle exactly zero", float(np.abs(a).max())))
            continue
        err = float(np.abs(np.asarray(a, dtype=np.float64) - r).max())
        den = float(np.abs(r).max())
        worst = max(worst, err / den)
        if err > tol * den + atol:
            bad.append((n, f"{err / den:.3e}"))
    assert not bad, f"{what}: {len(bad)}/{len(names)} tensors out of tolerance: {bad}"
    return worst


def _keras_adam_expected(w0, g, t, lr=2e-4, b1=0.5, b2=0.999, eps=1e-7, m=None, v=None):
    """Keras-Adam update in float64 from given gradients."""
    m = np.zeros_like(g, dtype=np.float64) if m is None else m
    v = np.zeros_like(g, dtype=np.float64) if v is None else v
    g = g.astype(np.float64)
    m = m + (g - m) * (1 - b1)
    v = v + (g * g - v) * (1 - b2)
    alpha = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return w0.astype(np.float64) - alpha * m / (np.sqrt(v) + eps), m, v


def _grad_errors(dev, ref64, ref32):
    """Per-tensor (device error, fp32-oracle error) relative to max|ref64|."""
    out = []
    for a, r64, r32 in zip(dev, ref64, ref32):
        r64 = r64.detach().numpy(); r32 = r32.detach().numpy().astype(np.float64)
        den = np.abs(r64).max()
        if den == 0.0:
            out.append((float(np.abs(a).max()), 0.0, True))
        else:
            out.append((float(np.abs(a - r64).max() / den), float(np.abs(r32 - r64).max() / den), False))
    return out


@pytest.mark.parametrize("batch,channels", [(1, 3), (2, 3), (2, 1), (3, 3)])
def test_fp32_step_matches_oracle(batch, channels):
    """fp32 path, one train step: losses <=1e-4 vs the float64 oracle; every gradient tensor within
    max(1e-4, GRAD_FACTOR x the float32 oracle's own deviation from float64) — at batch >= 2 the discriminator's
    real/fake gradients nearly cancel at initialisation and BatchNorm projects the conv gradients, so
    ANY float32 implementation (torch-CPU included) sits 1e-3..1e-2 from float64 there; and the
    Keras-Adam update reproduces the float64 formula applied to the device's own gradients."""
    m, g_np, d_np = _build("fp32", channels)
    x, y = _inputs(batch, 256, channels)
    names_g = [v.name for v in m.generator.trainable_variables]
    names_d = [v.name for v in m.discriminator.trainable_variables]
    masks = O.generator_keep_masks(SEED, m.ctx.call_counter(), 0, batch, 256)
    w0_g = m.generator.get_weights(); w0_d = m.discriminator.get_weights()
    losses = m.train_step(x, y, True)
    ref = {}
    for dt in (torch.float64, torch.float32):
        gp, dp = O.to_torch(g_np, dt), O.to_torch(d_np, dt)
        ref[dt] = O.pix2pix_losses_and_grads(gp, dp, torch.tensor(x, dtype=dt), torch.tensor(y, dtype=dt), 100.0, masks)
    for a, r in zip(losses, ref[torch.float64][0]):
        assert abs(float(a) - float(r)) <= 1e-4 * max(1.0, abs(float(r))), (list(map(float, losses)), ref[torch.float64][0])
    for tag, names, model, idx in (("dG", names_g, m.generator, 1), ("dD", names_d, m.discriminator, 2)):
        dev = [v.grad() for v in model.trainable_variables]
        errs = _grad_errors(dev, ref[torch.float64][idx], ref[torch.float32][idx])
        ratios = [e / max(e32, 1e-12) for e, e32, zero in errs if not zero and e > 1e-4]
        print(f"B={batch} C={channels} {tag}: worst dev err={max(e for e, _, _ in errs):.2e}, "
              f"worst dev/fp32-oracle ratio among tensors above 1e-4: {max(ratios) if ratios else 0:.2f}")
        bad = [(n, f"dev={e:.2e}", f"fp32-oracle={e32:.2e}") for n, (e, e32, zero) in zip(names, errs)
               if (zero and e >= 1e-10) or (not zero and e > max(1e-4, GRAD_FACTOR * e32))]
        assert not bad, f"{tag}: {len(bad)}/{len(names)} tensors out of tolerance: {bad}"
        # Keras-Adam arithmetic: exact (float32 rounding) given the device's own gradients
        w0 = w0_g if idx == 1 else w0_d
        for n, v, w_before, g in zip(names, model.trainable_variables, w0, dev):
            want, _, _ = _keras_adam_expected(w_before, g, 1)
            assert np.abs(v.numpy() - want).max() <= 2e-7 * max(1.0, np.abs(want).max()) + 1e-9, (tag, n)
    assert m.generator_optimizer.iterations == 1 and m.discriminator_optimizer.iterations == 1
    m.ctx.close()


def test_fp32_three_steps_track_oracle():
    """Free-running fp32 trajectory for N=3 steps at batch 1 (BASELINE config 1): the four losses
    stay within 1e-4 of the float64 oracle at every step."""
    m, g_np, d_np = _build("fp32", 3)
    gp, dp, go, do = _oracle_state(g_np, d_np)
    x, y = _inputs(1, 256, 3)
    xt, yt = torch.tensor(x, dtype=torch.float64), torch.tensor(y, dtype=torch.float64)
    names_g = [v.name for v in m.generator.trainable_variables]
    names_d = [v.name for v in m.discriminator.trainable_variables]
    for step in range(3):
        masks = O.generator_keep_masks(SEED, m.ctx.call_counter(), 0, 1, 256)
        losses = m.train_step(x, y, True)
        ref_losses, gg, dg = O.pix2pix_train_step(gp, dp, go, do, xt, yt, 100.0, True, masks)
        for a, r in zip(losses, ref_losses):
            assert abs(float(a) - r) <= 1e-4 * max(1.0, abs(r)), (step, list(map(float, losses)), ref_losses)
    # Weights are not compared tensor-by-tensor after several steps: Keras-Adam's early updates are
    # ~lr*sign(g), so any gradient below the fp32 noise floor separates that weight by 2*lr per step
    # (test_fp32_step_matches_oracle checks the update arithmetic exactly instead).  What must hold is
    # that the weights moved by no more than N*lr from the oracle's.
    for model, ref_p in ((m.generator, gp), (m.discriminator, dp)):
        for v, r in zip(model.trainable_variables, ref_p):
            assert np.abs(v.numpy() - r.detach().numpy()).max() <= 2 * 3 * 2e-4 * 1.01, v.name
    assert m.generator_optimizer.iterations == 3 and m.discriminator_optimizer.iterations == 3
    m.ctx.close()


def test_fp32_generator_forward_and_validation_step():
    m, g_np, d_np = _build("fp32", 3)
    gp, dp, go, do = _oracle_state(g_np, d_np)
    x, y = _inputs(2, 256, 3, seed=7)
    # model(x, training=True): predict semantics (pix2pix.py:228) — batch stats + dropout
    masks = O.generator_keep_masks(SEED, m.ctx.call_counter(), 0, 2, 256)
    out = m.generator(x, training=True)
    ref = O.generator_forward(gp, torch.tensor(x, dtype=torch.float64), "batchnorm", masks).detach().numpy()
    assert rel_err(out, ref) < 1e-4
    # discriminator([inp, tar], training=True)
    logits = m.discriminator([x, y], training=True)
    ref_l = O.discriminator_forward(dp, torch.tensor(x, dtype=torch.float64), torch.tensor(y, dtype=torch.float64)).detach().numpy()
    assert logits.shape == (2, 30, 30, 1) and rel_err(logits, ref_l) < 1e-4
    # train_step(..., training=False): same forward, no update, Adam t not advanced (pix2pix.py:208,292)
    w_before = m.generator.get_flat_params().copy()
    masks = O.generator_keep_masks(SEED, m.ctx.call_counter(), 0, 2, 256)
    losses = m.train_step(x, y, False)
    ref_losses, _, _ = O.pix2pix_train_step(gp, dp, go, do, torch.tensor(x, dtype=torch.float64),
                                            torch.tensor(y, dtype=torch.float64), 100.0, False, masks)
    for a, r in zip(losses, ref_losses):
        assert abs(float(a) - r) <= 1e-4 * max(1.0, abs(r))
    assert np.array_equal(w_before, m.generator.get_flat_params())
    assert m.generator_optimizer.iterations == 0
    # ragged tail batch (tf.data batch() without drop_remainder, pix2pix.py:163): B smaller than before
    masks = O.generator_keep_masks(SEED, m.ctx.call_counter(), 0, 1, 256)
    losses = m.train_step(x[:1], y[:1], False)
    ref_losses, _, _ = O.pix2pix_train_step(gp, dp, go, do, torch.tensor(x[:1], dtype=torch.float64),
                                            torch.tensor(y[:1], dtype=torch.float64), 100.0, False, masks)
    for a, r in zip(losses, ref_losses):
        assert abs(float(a) - r) <= 1e-4 * max(1.0, abs(r))
    m.ctx.close()


def test_fp32_intermediate_activations():
    """Per-layer localisation: raw conv outputs and activations of one generator forward."""
    m, g_np, d_np = _build("fp32", 3)
    gp = O.to_torch(g_np, torch.float64)
    x, _ = _inputs(2, 256, 3, seed=11)
    masks = O.generator_keep_masks(SEED, m.ctx.call_counter(), 0, 2, 256)
    m.generator(x)
    taps = {}
    O.generator_forward(gp, torch.tensor(x, dtype=torch.float64), "batchnorm", masks, taps=taps)
    for name in ["down1.a", "down2.z", "down2.a", "down5.a", "down8.a", "up1.a", "up3.a", "up4.z", "up7.a"]:
        dev = m.generator.debug_tensor(name)
        ref = taps[name].detach().numpy().reshape(-1)
        # down8 at batch 2 normalises over n=2 samples: x_hat = +-1/sqrt(1+eps/var) loses digits in fp32
        assert rel_err(dev, ref) < (3e-4 if name == "down8.a" else 1e-4), (name, rel_err(dev, ref))
    m.ctx.close()


def _cos(a, b):
    a = np.asarray(a, dtype=np.float64).ravel(); b = np.asarray(b, dtype=np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))


def test_bf16_step_tracks_oracle():
    """16-bit tcgen05 path (BASELINE.json: <=1e-2 relative on generator output and losses after N steps).
    Batch 8, N=3 train steps from identical weights, inputs and dropout masks (the benchmarked batch 64 and
    N=10 are in test_gpu_parity_configs.py):
      * step 0 (identical weights): generator output max-rel <= 1e-2 and L2 <= 5e-3, losses <= 1e-2,
        gradients point the same way (cosine >= 0.95 for every conv kernel; measured 0.969..1.000);
      * after every one of the N=3 steps the four losses stay within 1e-2 of the free-running float64
        oracle;
      * after N=3 steps the generator output is within 1e-2 of the oracle evaluated AT THE DEVICE'S
        WEIGHTS.  (The free-running generator outputs are printed, not asserted: Keras-Adam's first
        updates are +-lr for every weight whatever the gradient magnitude, so a rounding-induced
        sign flip of a near-zero gradient separates that weight by 2*lr = 2% of its init std; no
        reduced-precision implementation can follow the float64 trajectory weight by weight.)"""
    B = 8
    m, g_np, d_np = _build("bf16", 3)
    gp, dp, go, do = _oracle_state(g_np, d_np)
    x, y = _inputs(B, 256, 3)
    xt, yt = torch.tensor(x, dtype=torch.float64), torch.tensor(y, dtype=torch.float64)
    names_g = [v.name for v in m.generator.trainable_variables]
    # step 0: forward parity at identical weights
    masks = O.generator_keep_masks(SEED, m.ctx.call_counter(), 0, B, 256)
    out = m.generator(x)
    ref = O.generator_forward(gp, xt, "batchnorm", masks).detach().numpy()
    l2 = float(np.linalg.norm(out - ref) / np.linalg.norm(ref))
    print(f"bf16 step 0: gen_out l2_rel={l2:.3e} max_rel={rel_err(out, ref):.3e}")
    assert l2 < 5e-3 and rel_err(out, ref) < 1e-2       # SURVEY 8c metric (max-rel) and L2; fp16 activations: measured ~1e-3
    for step in range(3):
        masks = O.generator_keep_masks(SEED, m.ctx.call_counter(), 0, B, 256)
        losses = m.train_step(x, y, True)
        ref_losses, gg, dg = O.pix2pix_train_step(gp, dp, go, do, xt, yt, 100.0, True, masks)
        for a, r in zip(losses, ref_losses):
            assert abs(float(a) - r) <= 1e-2 * max(1.0, abs(r)), (step, list(map(float, losses)), ref_losses)
        if step == 0:
            cos = {n: _cos(v.grad(), g.numpy()) for n, v, g in zip(names_g, m.generator.trainable_variables, gg)
                   if n.endswith(".kernel")}
            print("bf16 step 0 gradient cosines:", {k: round(c, 4) for k, c in cos.items()})
            assert min(cos.values()) > 0.95, cos
    masks = O.generator_keep_masks(SEED, m.ctx.call_counter(), 0, B, 256)
    out = m.generator(x)
    ref_free = O.generator_forward(gp, xt, "batchnorm", masks).detach().numpy()
    gp_dev = O.to_torch(m.generator.get_weights(), torch.float64)
    ref_sync = O.generator_forward(gp_dev, xt, "batchnorm", masks).detach().numpy()
    l2_free = float(np.linalg.norm(out - ref_free) / np.linalg.norm(ref_free))
    l2_sync = float(np.linalg.norm(out - ref_sync) / np.linalg.norm(ref_sync))
    print(f"bf16 after 3 steps: gen_out l2_rel vs oracle at device weights={l2_sync:.3e}, vs free-running oracle={l2_free:.3e}")
    # the forward arithmetic at the device's own weights holds the stated tolerance with margin (fp16 activations)
    assert l2_sync < 5e-3 and rel_err(out, ref_sync) < 1e-2
    m.ctx.close()


def test_cuda_graph_replay_matches_eager():
    """gan_ctx_set_graphs(1): the captured step (replayed from the 3rd call on) must reproduce the
    eager step — same losses step after step (fresh dropout masks and Adam bias correction come from
    device-resident counters) and the same weights, up to fp32 atomic-accumulation order."""
    x, y = _inputs(2, 256, 3, seed=5)
    runs = []
    for graphs in (False, True):
        m, _, _ = _build("bf16", 3)
        m.ctx.set_graphs(graphs)
        losses = [[float(v) for v in m.train_step(x, y, True)] for _ in range(5)]
        val = [float(v) for v in m.train_step(x, y, False)]
        runs.append((losses, val, m.generator.get_flat_params(), m.ctx.call_counter(), m.generator_optimizer.iterations))
        m.ctx.close()
    (l0, v0, w0, c0, t0), (l1, v1, w1, c1, t1) = runs
    assert c0 == c1 == 6 and t0 == t1 == 5
    for a, b in zip(l0 + [v0], l1 + [v1]):
        for p, q in zip(a, b):
            assert abs(p - q) <= 2e-3 * max(1.0, abs(q)), (l0, l1)
    assert len({tuple(l) for l in l1}) == 5                     # every step differs: counters advance under replay
    assert np.abs(w0 - w1).max() <= 2 * 5 * 2e-4 * 1.01


def test_prefetch_is_transparent():
    """gan_ctx_prefetch (the reference's dataset.prefetch, pix2pix.py:163): copying the next batch on a
    copy stream while the current step runs must not change any result."""
    batches = [_inputs(2, 256, 3, seed=s) for s in (11, 12, 13)]
    res = []
    for use_prefetch in (False, True):
        m, _, _ = _build("fp32", 3)
        m.ctx.set_graphs(use_prefetch)        # exercise the graph path together with prefetch
        out = []
        if use_prefetch:
            m.ctx.prefetch(*batches[0])
        for i, (x, y) in enumerate(batches):
            losses = m.train_step(x, y, True, sync=False)
            if use_prefetch and i + 1 < len(batches):
                m.ctx.prefetch(*batches[i + 1])
            out.append([float(v) for v in m.ctx.last_losses(4)])
        res.append(out)
        m.ctx.close()
    for a, b in zip(*res):
        for p, q in zip(a, b):
            assert abs(p - q) <= 1e-4 * max(1.0, abs(q)), res
