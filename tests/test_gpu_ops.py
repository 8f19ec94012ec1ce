"""GPU parity of every convolution kernel family against the CPU oracle, through the C-ABI
(gan_op_conv): Conv2D 4x4 s2 'same', ZeroPad+Conv2D 4x4 s1, Conv2DTranspose 4x4 s2 'same', each in
forward / data-gradient / weight-gradient form, on the FFMA (fp32, 16-bit) and tcgen05 engines.

16-bit mode storage (common.cuh): fp16 for activations, gradients and both weight packs (tcgen05 kind::f16 needs
A and B in the same format: an f16 x bf16 MMA faults as an illegal instruction on B200).  The oracle is fed
operands rounded the same way, so what is left is accumulation order and the rounding of the stored result.
Both storage formats are run (GAN_B200_ACT=bf16 is round 1's all-bf16 storage)."""
import os

import numpy as np
import pytest

from helpers import oracle_conv, oracle_conv_grads, rel_err, act_round

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4      # BASELINE.json: <=1e-4 relative in the fp32 path
BF16_TOL = 1e-2      # BASELINE.json: <=1e-2 relative in the bf16 path


@pytest.fixture(scope="module")
def ctx32():
    from gan_b200 import Context
    c = Context(0, "fp32", 1)
    yield c
    c.close()


@pytest.fixture(scope="module", params=["f16", "bf16"])
def ctx16(request):
    from gan_b200 import Context
    old = os.environ.get("GAN_B200_ACT")
    os.environ["GAN_B200_ACT"] = request.param          # read when the context is created
    c = Context(0, "bf16", 1)
    c.act = request.param
    if old is None:
        del os.environ["GAN_B200_ACT"]
    else:
        os.environ["GAN_B200_ACT"] = old
    yield c
    c.close()


def _refs(ctx, kind, x, wt, dy):
    """Oracle results on operands rounded the way the device stores them."""
    a = ctx.act
    y_ref = oracle_conv(kind, act_round(x, a), act_round(wt, a)).numpy()
    dx_ref, dw_ref = oracle_conv_grads(kind, act_round(x, a), act_round(wt, a), act_round(dy, a))
    return y_ref, dx_ref, dw_ref


def _out_tol(ctx):
    return 6e-3 if ctx.act == "bf16" else 1e-3          # rounding of the stored forward result (2^-9 vs 2^-12 relative)


def _case(kind, b, h, w, cin, cout, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, size=(b, h, w, cin)).astype(np.float32)
    wshape = (4, 4, cout, cin) if kind == 2 else (4, 4, cin, cout)
    wt = rng.normal(0, 0.05, size=wshape).astype(np.float32)
    ho, wo = {0: (h // 2, w // 2), 1: (h - 1, w - 1), 2: (2 * h, 2 * w)}[kind]
    dy = rng.normal(0, 1, size=(b, ho, wo, cout)).astype(np.float32)
    return x, wt, dy


SMALL = [  # kind, B, H, W, Cin, Cout   (ragged channel counts hit the scalar + skinny kernels)
    (0, 2, 8, 8, 3, 64), (0, 1, 16, 16, 6, 64), (0, 2, 8, 8, 64, 128), (0, 3, 4, 4, 16, 20), (0, 1, 2, 2, 64, 64),
    (1, 2, 8, 8, 64, 1), (1, 1, 9, 7, 32, 64), (1, 2, 6, 6, 8, 3),
    (2, 2, 4, 4, 64, 64), (2, 1, 8, 8, 128, 3), (2, 2, 1, 1, 32, 64), (2, 1, 4, 4, 64, 1), (2, 2, 3, 5, 12, 20),
]


@pytest.mark.parametrize("kind,b,h,w,cin,cout", SMALL)
def test_ffma_fp32_forward_dgrad_wgrad(ctx32, kind, b, h, w, cin, cout):
    x, wt, dy = _case(kind, b, h, w, cin, cout)
    y_ref = oracle_conv(kind, x, wt).numpy()
    dx_ref, dw_ref = oracle_conv_grads(kind, x, wt, dy)
    y = ctx32.op_conv(kind, 0, x, wt, b, h, w, cin, cout, engine=0)
    dx = ctx32.op_conv(kind, 1, dy, wt, b, h, w, cin, cout, engine=0)
    dw = ctx32.op_conv(kind, 2, x, dy, b, h, w, cin, cout, engine=0)
    assert rel_err(y, y_ref) < FP32_TOL
    assert rel_err(dx, dx_ref) < FP32_TOL
    assert rel_err(dw, dw_ref) < FP32_TOL


@pytest.mark.parametrize("kind,b,h,w,cin,cout", SMALL[:3] + SMALL[5:6] + SMALL[8:10])
def test_ffma_bf16_forward_dgrad_wgrad(ctx16, kind, b, h, w, cin, cout):
    x, wt, dy = _case(kind, b, h, w, cin, cout)
    y_ref, dx_ref, dw_ref = _refs(ctx16, kind, x, wt, dy)          # device sees rounded operands
    y = ctx16.op_conv(kind, 0, x, wt, b, h, w, cin, cout, engine=0)
    dx = ctx16.op_conv(kind, 1, dy, wt, b, h, w, cin, cout, engine=0)
    dw = ctx16.op_conv(kind, 2, x, dy, b, h, w, cin, cout, engine=0)
    assert rel_err(y, y_ref) < BF16_TOL
    assert rel_err(dx, dx_ref) < BF16_TOL
    assert rel_err(dw, dw_ref) < 1e-4        # wgrad output is fp32: only accumulation-order error


UMMA_CASES = [  # shapes where Kc and Nc are multiples of 64 (the tcgen05 tile constraints)
    (0, 2, 16, 16, 64, 128), (0, 1, 32, 32, 128, 64), (0, 4, 4, 4, 64, 64), (0, 3, 2, 2, 128, 128),
    (1, 2, 8, 8, 64, 128), (1, 1, 32, 32, 64, 64), (1, 2, 9, 11, 64, 64),
    (2, 2, 8, 8, 128, 64), (2, 1, 16, 16, 64, 128), (2, 5, 2, 2, 64, 64), (2, 2, 1, 1, 128, 128),
    (1, 20, 32, 32, 64, 256), (0, 19, 64, 64, 64, 256),     # >= 148 tiles with Cout % 256 == 0: N=256 tile variant
]


@pytest.mark.parametrize("kind,b,h,w,cin,cout", UMMA_CASES)
def test_umma_forward_dgrad_match_oracle(ctx16, kind, b, h, w, cin, cout):
    x, wt, dy = _case(kind, b, h, w, cin, cout, seed=3)
    y_ref, dx_ref, _ = _refs(ctx16, kind, x, wt, dy)
    y = ctx16.op_conv(kind, 0, x, wt, b, h, w, cin, cout, engine=1)
    dx = ctx16.op_conv(kind, 1, dy, wt, b, h, w, cin, cout, engine=1)
    # rounding of the stored output is the only error source besides fp32 accumulation order
    assert rel_err(y, y_ref) < _out_tol(ctx16)
    assert rel_err(dx, dx_ref) < _out_tol(ctx16)
    # and the two engines agree to output rounding
    y_f = ctx16.op_conv(kind, 0, x, wt, b, h, w, cin, cout, engine=0)
    assert rel_err(y, y_f) < _out_tol(ctx16)


@pytest.mark.parametrize("kind,b,h,w,cin,cout", UMMA_CASES)
def test_umma_wgrad_matches_oracle(ctx16, kind, b, h, w, cin, cout):
    x, wt, dy = _case(kind, b, h, w, cin, cout, seed=4)
    _, _, dw_ref = _refs(ctx16, kind, x, wt, dy)
    dw = ctx16.op_conv(kind, 2, x, dy, b, h, w, cin, cout, engine=1)
    assert rel_err(dw, dw_ref) < 1e-4        # 16-bit operands, fp32 accumulate: exact products


SMALLC = [  # first layers (Cin in {1,3,6}) and heads (Cout in {1,3}): channel-padded tcgen05 paths
    (0, 2, 16, 16, 3, 64), (0, 1, 32, 32, 6, 64), (0, 2, 16, 16, 1, 64),
    (1, 2, 8, 8, 64, 1), (1, 1, 16, 16, 512, 1), (1, 2, 31, 31, 128, 1),
    (2, 2, 8, 8, 128, 3), (2, 1, 8, 8, 128, 1), (2, 3, 16, 16, 64, 3),
]


@pytest.mark.parametrize("kind,b,h,w,cin,cout", SMALLC)
def test_umma_small_channel_layers(ctx16, kind, b, h, w, cin, cout):
    x, wt, dy = _case(kind, b, h, w, cin, cout, seed=6)
    y_ref, dx_ref, dw_ref = _refs(ctx16, kind, x, wt, dy)
    y = ctx16.op_conv(kind, 0, x, wt, b, h, w, cin, cout, engine=1)
    dx = ctx16.op_conv(kind, 1, dy, wt, b, h, w, cin, cout, engine=1)
    dw = ctx16.op_conv(kind, 2, x, dy, b, h, w, cin, cout, engine=1)
    assert rel_err(y, y_ref) < _out_tol(ctx16)
    assert rel_err(dx, dx_ref) < _out_tol(ctx16)
    assert rel_err(dw, dw_ref) < 1e-4
