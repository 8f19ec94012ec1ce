"""What can ANY 16-bit tensor-core implementation of the generator hold?  (CPU only, test infrastructure.)

The float64 oracle forward is re-run with rounding injected exactly where a tcgen05 path must round:
the MMA operands (activations and weights), optionally also the raw conv output z before the
normalisation (what the device stores between the conv and the norm kernel).  Everything else —
accumulation, statistics, normalisation, activation — stays float64, i.e. this is the FLOOR for an
implementation with that operand format, not a model of our kernels.

    python scripts/precision_floor.py [B ...]

Prints max-rel and L2 of the generator output against the unquantised float64 oracle (SURVEY 8c metrics).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gan_oracle as O  # noqa: E402

SEED = 123


def rounder(dt):
    if dt is None:
        return lambda t: t
    return lambda t: t.to(dt).to(torch.float64)


def run(B, op_dt, z_dt, fp32_z_small=False):
    qo, qz = rounder(op_dt), rounder(z_dt)
    orig = (O.conv2d_s2_same, O.conv2d_transpose_s2_same)

    def conv(x, w, _f=orig[0]):
        z = _f(qo(x), qo(w))
        small = fp32_z_small and z.shape[2] <= 8
        return z if small else qz(z)

    def convT(x, f, bias=None, _f=orig[1]):
        z = _f(qo(x), qo(f), bias)
        small = fp32_z_small and z.shape[2] <= 8
        return z if (bias is not None or small) else qz(z)

    rng = np.random.default_rng(SEED + 1)
    g_np = O.init_params(O.generator_spec(3), rng, "batchnorm")
    gp = O.to_torch(g_np, torch.float64, False)
    x = torch.tensor(O.synthetic_images(np.random.default_rng(SEED), B, 256, 256, 3), dtype=torch.float64)
    masks = O.generator_keep_masks(SEED, 0, 0, B, 256)
    with torch.no_grad():
        ref = O.generator_forward(gp, x, "batchnorm", masks).numpy()
        O.conv2d_s2_same, O.conv2d_transpose_s2_same = conv, convT
        try:
            out = O.generator_forward(gp, x, "batchnorm", masks).numpy()
        finally:
            O.conv2d_s2_same, O.conv2d_transpose_s2_same = orig
    d = out - ref
    return np.abs(d).max() / np.abs(ref).max(), np.linalg.norm(d) / np.linalg.norm(ref)


if __name__ == "__main__":
    torch.set_num_threads(4)
    for B in [int(a) for a in sys.argv[1:]] or [2, 8]:
        for name, op, z, small in (("bf16 operands, exact z (floor of any bf16 path)", torch.bfloat16, None, False),
                                   ("bf16 operands, bf16 z (what round 1 stored)", torch.bfloat16, torch.bfloat16, False),
                                   ("bf16 operands, bf16 z, fp32 z on <=8x8 layers", torch.bfloat16, torch.bfloat16, True),
                                   ("fp16 operands, exact z", torch.float16, None, False),
                                   ("fp16 operands, fp16 z", torch.float16, torch.float16, False)):
            mx, l2 = run(B, op, z, small)
            print(f"B={B:3d}  {name:50s} max-rel={mx:.3e}  L2={l2:.3e}", flush=True)
