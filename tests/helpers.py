"""Shared helpers for the parity tests (oracle side). The oracle is the checker only."""
import numpy as np
import torch

from oracle import gan_oracle as O

KINDS = {0: "conv_s2", 1: "conv_s1p", 2: "convT_s2"}


def rel_err(a, b):
    """max|a-b| / max|b| — per-tensor relative error as defined in SURVEY 8c."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def oracle_conv(kind, x_nhwc, w):
    x = torch.tensor(x_nhwc, dtype=torch.float64).permute(0, 3, 1, 2)
    w = torch.tensor(w, dtype=torch.float64)
    if kind == 0:
        y = O.conv2d_s2_same(x, w)
    elif kind == 1:
        y = O.conv2d_s1_pad1(x, w)
    else:
        y = O.conv2d_transpose_s2_same(x, w)
    return y.permute(0, 2, 3, 1).contiguous()


def oracle_conv_grads(kind, x_nhwc, w, dy_nhwc):
    """Returns (dx NHWC, dw in TF layout) of sum(y*dy)."""
    x = torch.tensor(x_nhwc, dtype=torch.float64, requires_grad=True)
    wt = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    xc = x.permute(0, 3, 1, 2)
    if kind == 0:
        y = O.conv2d_s2_same(xc, wt)
    elif kind == 1:
        y = O.conv2d_s1_pad1(xc, wt)
    else:
        y = O.conv2d_transpose_s2_same(xc, wt)
    y = y.permute(0, 2, 3, 1)
    (y * torch.tensor(dy_nhwc, dtype=torch.float64)).sum().backward()
    return x.grad.numpy(), wt.grad.numpy()


def bf16_round(a):
    return torch.tensor(np.asarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def f16_round(a):
    return torch.tensor(np.asarray(a, dtype=np.float32)).to(torch.float16).to(torch.float32).numpy()


def act_round(a, act="f16"):
    """Rounding the device applies to the stored operands (activations, gradients, weight packs) in the 16-bit mode:
    fp16 by default, bf16 under GAN_B200_ACT=bf16 (common.cuh)."""
    return f16_round(a) if act == "f16" else bf16_round(a)


def make_pix2pix(seed_w, channels, dtype):
    """Oracle-side Pix2Pix parameters (numpy lists) from default_rng(seed_w) in Keras variable order."""
    rng = np.random.default_rng(seed_w)
    g = O.init_params(O.generator_spec(channels), rng, "batchnorm")
    d = O.init_params(O.discriminator_spec(channels, True), rng, "batchnorm")
    return g, d


def load_model(model, arrays):
    for v, a in zip(model.trainable_variables, arrays):
        assert tuple(v.shape) == tuple(a.shape), (v.name, v.shape, a.shape)
        v.assign(a)
