"""CPU tests of the oracle: direct-definition cross-checks, known answers, golden fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import gan_oracle as O, direct as D

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _t(a):
    return torch.tensor(a, dtype=torch.float64).permute(0, 3, 1, 2)


def _n(t):
    return t.permute(0, 2, 3, 1).numpy()


@pytest.mark.parametrize("shape", [(2, 8, 8, 3, 5), (1, 4, 6, 2, 7), (2, 2, 2, 4, 4)])
def test_conv_layout_conventions_against_direct_definition(shape):
    n, h, w, ci, co = shape
    rng = np.random.default_rng(0)
    x = rng.normal(size=(n, h, w, ci)); k = rng.normal(size=(4, 4, ci, co)); f = rng.normal(size=(4, 4, co, ci))
    b = rng.normal(size=(co,))
    assert np.abs(_n(O.conv2d_s2_same(_t(x), torch.tensor(k))) - D.conv2d_s2_same(x, k)).max() < 1e-12
    assert np.abs(_n(O.conv2d_s1_pad1(_t(x), torch.tensor(k), torch.tensor(b))) - D.conv2d_s1_pad1(x, k, b)).max() < 1e-12
    yt = D.conv2d_transpose_s2_same(x, f, b)
    assert np.abs(_n(O.conv2d_transpose_s2_same(_t(x), torch.tensor(f), torch.tensor(b))) - yt).max() < 1e-12
    # the parity-class (gather) form the GPU kernels use equals the scatter definition
    assert np.abs(D.conv2d_transpose_s2_same_gather(x, f) + b - yt).max() < 1e-12


def test_transposed_conv_is_the_input_gradient_of_conv():
    """App. A.3: Conv2DTranspose(f) == d/dx of Conv2D with kernel f read as (kh,kw,in=co,out=ci)."""
    rng = np.random.default_rng(1)
    x = torch.tensor(rng.normal(size=(1, 3, 4, 4)), requires_grad=False)          # NCHW, 3 ch
    f = torch.tensor(rng.normal(size=(4, 4, 5, 3)))                               # (kh,kw,out=5,in=3)
    y = O.conv2d_transpose_s2_same(x, f)                                          # (1,5,8,8)
    z = torch.zeros(1, 5, 8, 8, dtype=torch.float64, requires_grad=True)
    out = O.conv2d_s2_same(z, f)                                                  # kernel (kh,kw,in=5,out=3)
    (g,) = torch.autograd.grad((out * x).sum(), z)
    assert (g - y).abs().max() < 1e-12


def test_norms_losses_adam_against_direct_definition():
    rng = np.random.default_rng(2)
    x = rng.normal(size=(3, 5, 4, 6)); g = rng.normal(size=(6,)); b = rng.normal(size=(6,))
    y, _, _ = O.batch_norm_train(_t(x), torch.tensor(g), torch.tensor(b))
    assert np.abs(_n(y) - D.batch_norm_train(x, g, b)).max() < 1e-12
    assert np.abs(_n(O.instance_norm(_t(x), torch.tensor(g), torch.tensor(b))) - D.instance_norm(x, g, b)).max() < 1e-12
    lg = rng.normal(size=(2, 30, 30, 1)) * 3
    for z in (0.0, 1.0):
        assert abs(float(O.bce_from_logits(torch.tensor(lg), z)) - D.bce_from_logits(lg, z)) < 1e-12
    p = torch.tensor(rng.normal(size=(7,)), requires_grad=True)
    opt = O.KerasAdam([p])
    th, m, v = p.detach().numpy().copy(), np.zeros(7), np.zeros(7)
    for t in range(1, 4):
        gr = rng.normal(size=(7,))
        opt.apply_gradients([torch.tensor(gr)], [p])
        th, m, v = D.keras_adam_step(th, gr, m, v, t)
        assert np.abs(p.detach().numpy() - th).max() < 1e-15


def test_degenerate_bottleneck_is_exactly_beta():
    """BatchNorm with n=1 and InstanceNorm on 1x1 give exactly beta."""
    x = torch.randn(1, 512, 1, 1, dtype=torch.float64)
    g = torch.ones(512, dtype=torch.float64); b = torch.full((512,), 0.25, dtype=torch.float64)
    y, _, _ = O.batch_norm_train(x, g, b)
    assert torch.equal(y.flatten(), b)
    x4 = torch.randn(4, 512, 1, 1, dtype=torch.float64)
    assert torch.equal(O.instance_norm(x4, g, b), b.view(1, -1, 1, 1).expand(4, -1, 1, 1))


def test_parameter_counts_and_variable_order():
    assert O.num_params(O.generator_spec(3)) == 54_414_979
    assert O.num_params(O.generator_spec(1)) == 54_408_833
    assert O.num_params(O.discriminator_spec(3, True)) == 2_768_641
    assert O.num_params(O.discriminator_spec(3, False)) == 2_765_569
    names = [n for n, _, _ in O.generator_spec(3)]
    assert len(names) == 45 and names[0] == "down1.kernel" and names[1] == "down2.kernel"
    assert names[-2:] == ["last.kernel", "last.bias"] and names[22] == "up1.kernel"
    assert len(O.discriminator_spec(3, True)) == 12


def test_philox_published_known_answers():
    """Random123 kat_vectors for philox4x32-10 — an external pin of the dropout RNG."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        w = O.philox4x32_10(*[np.array([v], dtype=np.uint64) for v in c], k[0], k[1])
        assert tuple(int(x[0]) for x in w) == want
    g = np.load(os.path.join(GOLD, "philox_kat.npz"))
    c = np.arange(8, dtype=np.uint64)
    w = O.philox4x32_10(c, c + 1, c + 2, c + 3, 0xDEADBEEF, 0x12345678)
    for i in range(4):
        assert np.array_equal(w[i], g[f"w{i}"])
    m = O.dropout_keep_mask(123, 0, 1, 0, (2, 4, 4, 512))
    assert 0.45 < m.mean() < 0.55
    # keyed on the global sample index: rank-local slices reproduce the unsharded mask
    assert np.array_equal(O.dropout_keep_mask(123, 0, 1, 1, (1, 4, 4, 512))[0], m[1])


def test_pix2pix_golden_fixture():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    got = mg.pix2pix_case(1, 3, 2)
    want = np.load(os.path.join(GOLD, "pix2pix_b1_c3.npz"))
    for k in want.files:
        np.testing.assert_allclose(got[k], want[k], rtol=1e-9, atol=1e-12, err_msg=k)
    # B=1 at 256^2: the down8 BatchNorm sees n=1 -> zero gradient for down8's kernel/gamma
    names = [n for n, _, _ in O.generator_spec(3)]
    assert want["g_grad_norms_0"][names.index("down8.kernel")] == 0.0
    assert want["g_grad_norms_0"][names.index("down7.kernel")] > 0.0


def test_float32_oracle_agrees_with_float64():
    rng = np.random.default_rng(124)
    g_np = O.init_params(O.generator_spec(3), rng, "batchnorm")
    d_np = O.init_params(O.discriminator_spec(3, True), rng, "batchnorm")
    irng = np.random.default_rng(123)
    x = O.synthetic_images(irng, 1, 256, 256, 3); y = O.synthetic_images(irng, 1, 256, 256, 3)
    res = []
    for dt in (torch.float32, torch.float64):
        gp, dp = O.to_torch(g_np, dt), O.to_torch(d_np, dt)
        masks = O.generator_keep_masks(123, 0, 0, 1, 256)
        losses, _, _, _ = O.pix2pix_losses_and_grads(gp, dp, torch.tensor(x, dtype=dt), torch.tensor(y, dtype=dt),
                                                     100.0, masks, want_grads=False)
        res.append([float(l) for l in losses])
    for a, b in zip(*res):
        assert abs(a - b) <= 1e-5 * max(1.0, abs(b))


def test_cyclegan_loss_identities_in_the_golden():
    """total_gen_g = gen_g + total_cycle + identity_y (cycle_gan.py:243-244): consistency of the stored losses."""
    want = np.load(os.path.join(GOLD, "cyclegan_b1_c3.npz"))
    assert want["losses_0"].shape == (7,)
    l = want["losses_0"]
    assert l[3] >= l[0] + l[2] - 1e-12 and l[4] >= l[1] + l[2] - 1e-12


def test_cyclegan_single_sweep_equals_four_tape_gradients():
    """ONE backward sweep of gen_g + gen_f + total_cycle + id_x + id_y (what the device runs) yields
    exactly the reference's generator_g and generator_f gradients, which TensorFlow computes with two separate
    tape.gradient calls on total_gen_g / total_gen_f (cycle_gan.py:252-255): d gen_f / dG = d id_x / dG = 0 and
    d gen_g / dF = d id_y / dF = 0, and total_cycle is shared."""
    rng = np.random.default_rng(5)
    specs = [O.generator_spec(3), O.generator_spec(3), O.discriminator_spec(3, False), O.discriminator_spec(3, False)]
    nets = [O.to_torch(O.init_params(s, rng, "instancenorm"), torch.float64) for s in specs]
    irng = np.random.default_rng(6)
    x = torch.tensor(O.synthetic_images(irng, 1, 256, 256, 3), dtype=torch.float64)
    y = torch.tensor(O.synthetic_images(irng, 1, 256, 256, 3), dtype=torch.float64)
    losses, _, _, _, _, outs = O.cyclegan_losses_and_grads(nets[0], nets[1], nets[2], nets[3], x, y, 10.0, None, want_grads=False)
    gen_g, gen_f, total_cycle, total_g, total_f = losses[:5]
    g_g = torch.autograd.grad(total_g, nets[0], retain_graph=True)      # cycle_gan.py:252
    g_f = torch.autograd.grad(total_f, nets[1], retain_graph=True)      # cycle_gan.py:254
    id_y = total_g - gen_g - total_cycle
    id_x = total_f - gen_f - total_cycle
    single = gen_g + gen_f + total_cycle + id_x + id_y
    sweep_g = torch.autograd.grad(single, nets[0], retain_graph=True)
    sweep_f = torch.autograd.grad(single, nets[1])
    for a, b in list(zip(sweep_g, g_g)) + list(zip(sweep_f, g_f)):
        den = float(b.abs().max())
        assert float((a - b).abs().max()) <= 1e-12 * max(den, 1e-30) + 1e-18


@pytest.mark.parametrize("n,h,w,c,nsrc", [(2, 8, 8, 3, 1), (1, 8, 12, 3, 2), (2, 4, 8, 1, 2)])
def test_first_layer_operand_layouts_against_autograd(n, h, w, c, nsrc):
    """The slot-4 row / weight-pack / per-tap-cols index maps the first-layer kernels use (restated in oracle/direct.py)
    reproduce Conv2D 4x4 s2 'same' of the concatenated sources, its weight gradient and its input gradient."""
    rng = np.random.default_rng(7)
    srcs = [rng.normal(size=(n, h, w, c)) for _ in range(nsrc)]
    k = rng.normal(size=(4, 4, c * nsrc, 64))
    x = torch.tensor(np.concatenate(srcs, -1)).permute(0, 3, 1, 2).requires_grad_(True)
    kt = torch.tensor(k, requires_grad=True)
    y = O.conv2d_s2_same(x, kt)
    dz = rng.normal(size=(n, h // 2, w // 2, 64))
    (gx, gk) = torch.autograd.grad((y * _t(dz)).sum(), (x, kt))
    assert np.abs(D.conv2d_s2_same_via_rows(srcs, k) - _n(y.detach())).max() < 1e-11
    assert np.abs(D.conv2d_s2_same_wgrad_via_rows(srcs, dz) - gk.numpy()).max() < 1e-10
    c0 = c * (nsrc - 1)                                   # the generated image is the LAST source of concatenate([inp, tar])
    assert np.abs(D.conv2d_s2_same_dgrad_via_cols(dz, k, c0, c) - _n(gx)[..., c0:c0 + c]).max() < 1e-10
