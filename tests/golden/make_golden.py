"""Generates the golden fixtures under tests/golden/ from the CPU oracle (float64).

PARITY UNPINNED against TensorFlow (the reference cannot be imported here — see
oracle/gan_oracle.py); these fixtures pin the ORACLE against regressions: losses, a strided
sample of the generator output, per-tensor gradient L2 norms and post-Adam weight checksums for
small seeded runs.  Weights (57 M floats) are regenerated from the seed, never stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import gan_oracle as O  # noqa: E402

SEED = 123


def pix2pix_case(batch, channels, steps, size=256):
    rng = np.random.default_rng(SEED + 1)
    gp = O.to_torch(O.init_params(O.generator_spec(channels), rng, "batchnorm"))
    dp = O.to_torch(O.init_params(O.discriminator_spec(channels, True), rng, "batchnorm"))
    go, do = O.KerasAdam(gp), O.KerasAdam(dp)
    irng = np.random.default_rng(SEED)
    x = torch.tensor(O.synthetic_images(irng, batch, size, size, channels), dtype=torch.float64)
    y = torch.tensor(O.synthetic_images(irng, batch, size, size, channels), dtype=torch.float64)
    out = {}
    for s in range(steps):
        masks = O.generator_keep_masks(SEED, s, 0, batch, size)
        losses, gg, dg = O.pix2pix_train_step(gp, dp, go, do, x, y, 100.0, True, masks)
        out[f"losses_{s}"] = np.array(losses)
        out[f"g_grad_norms_{s}"] = np.array([float(g.norm()) for g in gg])
        out[f"d_grad_norms_{s}"] = np.array([float(g.norm()) for g in dg])
        out[f"g_weight_sums_{s}"] = np.array([float(p.detach().sum()) for p in gp])
        out[f"d_weight_sums_{s}"] = np.array([float(p.detach().sum()) for p in dp])
    masks = O.generator_keep_masks(SEED, steps, 0, batch, size)
    gen = O.generator_forward(gp, x, "batchnorm", masks).detach().numpy().reshape(-1)
    out["gen_sample"] = gen[::4099][:96]
    return out


def cyclegan_case(batch, channels, size=256):
    rng = np.random.default_rng(SEED + 1)
    specs = [O.generator_spec(channels), O.generator_spec(channels), O.discriminator_spec(channels, False),
             O.discriminator_spec(channels, False)]
    nets = [O.to_torch(O.init_params(s, rng, "instancenorm")) for s in specs]
    opts = [O.KerasAdam(p) for p in nets]
    irng = np.random.default_rng(SEED)
    x = torch.tensor(O.synthetic_images(irng, batch, size, size, channels), dtype=torch.float64)
    y = torch.tensor(O.synthetic_images(irng, batch, size, size, channels), dtype=torch.float64)
    calls = ['fake_y', 'cycled_x', 'fake_x', 'cycled_y', 'same_x', 'same_y']
    masks = {n: O.generator_keep_masks(SEED, i, 0, batch, size) for i, n in enumerate(calls)}
    losses, grads = O.cyclegan_train_step(nets, opts, x, y, 10.0, True, masks)
    out = {"losses_0": np.array(losses)}
    for tag, g, p in zip("gfxy", grads, nets):
        out[f"{tag}_grad_norms_0"] = np.array([float(t.norm()) for t in g])
        out[f"{tag}_weight_sums_0"] = np.array([float(t.detach().sum()) for t in p])
    return out


if __name__ == "__main__" and "--pipeline" not in sys.argv:     # `--pipeline`: only the input-pipeline fixture
    np.savez(os.path.join(HERE, "pix2pix_b1_c3.npz"), **pix2pix_case(1, 3, 2))
    np.savez(os.path.join(HERE, "pix2pix_b2_c1.npz"), **pix2pix_case(2, 1, 1))
    np.savez(os.path.join(HERE, "cyclegan_b1_c3.npz"), **cyclegan_case(1, 3))
    # Philox known answers (counter, key) -> first word, pinned for the CUDA/NumPy shared definition
    c = np.arange(8, dtype=np.uint64)
    w = O.philox4x32_10(c, c + 1, c + 2, c + 3, 0xDEADBEEF, 0x12345678)
    np.savez(os.path.join(HERE, "philox_kat.npz"), w0=w[0], w1=w[1], w2=w[2], w3=w[3])
    print("golden written")


def make_pipeline_golden():
    """Input-pipeline fixture (oracle/pipeline_oracle.py): a 37x90x3 pair and a 50x41x3 image through
    the four per-image paths at img_size 16."""
    from oracle import pipeline_oracle as P
    rng = np.random.default_rng(2024)
    pair = rng.integers(0, 256, size=(37, 90, 3), dtype=np.uint8)
    image = rng.integers(0, 256, size=(50, 41, 3), dtype=np.uint8)
    cy, cx, flip = 11, 29, True
    a, b = P.pix2pix_process_train(pair, 'left', 16, cy, cx, flip)
    pa, pb = P.pix2pix_process_pred(pair, 'right', 16)
    np.savez_compressed(os.path.join(HERE, "pipeline_small.npz"), pair=pair, image=image, cy=cy, cx=cx, flip=flip,
                        p2p_train_a=a, p2p_train_b=b, p2p_pred_a=pa, p2p_pred_b=pb,
                        cyc_train=P.cyclegan_process_train(image, 16, cy, cx, flip),
                        cyc_pred=P.cyclegan_process_pred(image, 16))


if __name__ == "__main__" and "--pipeline" in sys.argv:
    make_pipeline_golden()
