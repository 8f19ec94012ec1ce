"""Oracle vs the REAL reference (TensorFlow 2.6) — runs only when someone with a TF-2.6 environment has produced
``tests/golden/tf/*.npz`` with ``scripts/dump_tf_goldens.py`` (TensorFlow cannot be installed in the build container, so
the directory is empty there and this test is skipped: PARITY UNPINNED).  With a file present it pins
every TF-semantics item of SURVEY App. A: losses, all gradients, post-Adam weights and slots per step, float32
tolerance (TensorFlow computes in float32; the oracle in float64)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import gan_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "tf", "pix2pix_*.npz")))


def _rel(a, b):
    den = np.abs(b).max()
    return float(np.abs(np.asarray(a, np.float64) - b).max() / (den if den > 0 else 1.0))


@pytest.mark.skipif(not FILES, reason="no TensorFlow goldens (run scripts/dump_tf_goldens.py under TF 2.6)")
@pytest.mark.parametrize("path", FILES or ["-"])
def test_oracle_matches_tensorflow_pix2pix(path):
    d = np.load(path)
    seed, B, S, C, steps = (int(v) for v in d["protocol"])
    x = torch.tensor(d["x"], dtype=torch.float64); y = torch.tensor(d["y"], dtype=torch.float64)
    ng, nd = len(d["g/names"]), len(d["d/names"])
    # Keras variable order (App. A.8) must be what the oracle's spec assumes: shapes line up tensor by tensor
    g_np = [d[f"g/w0/{i}"] for i in range(ng)]; d_np = [d[f"d/w0/{i}"] for i in range(nd)]
    spec_g = O.init_params(O.generator_spec(C), np.random.default_rng(0), "batchnorm")
    spec_d = O.init_params(O.discriminator_spec(C, True), np.random.default_rng(0), "batchnorm")
    assert [a.shape for a in g_np] == [a.shape for a in spec_g] and [a.shape for a in d_np] == [a.shape for a in spec_d]
    gp, dp = O.to_torch(g_np, torch.float64), O.to_torch(d_np, torch.float64)
    go, do = O.KerasAdam(gp), O.KerasAdam(dp)
    for s in range(steps):
        masks = [torch.from_numpy(d[f"step{s}/mask{t}"].astype(np.float32)) for t in (1, 2, 3)]
        # the stored masks ARE the oracle's Philox masks for (seed, call=s): the dump and the oracle agree on the key
        want = O.generator_keep_masks(seed, s, 0, B, S)
        assert all(np.array_equal(a.numpy(), b.numpy()) for a, b in zip(masks, want))
        losses, gg, dg = O.pix2pix_train_step(gp, dp, go, do, x, y, 100.0, True, masks)
        for a, r in zip(losses, d[f"step{s}/losses"]):
            assert abs(a - r) <= 1e-4 * max(1.0, abs(r)), (s, losses, d[f"step{s}/losses"])
        for tag, grads, n in (("g", gg, ng), ("d", dg, nd)):
            for i in range(n):
                ref = d[f"step{s}/{tag}/grad/{i}"].astype(np.float64)
                if np.abs(ref).max() == 0:
                    assert np.abs(grads[i].numpy()).max() < 1e-10
                else:
                    # TensorFlow's float32 gradients are ill-conditioned at batch >= 2: loose bound here,
                    # the tight one is on the losses and on the batch-1 file
                    assert _rel(grads[i].numpy(), ref) <= (1e-4 if B == 1 else 1e-1), (s, tag, i)
        if B == 1:
            for tag, params, n in (("g", gp, ng), ("d", dp, nd)):
                for i in range(n):
                    assert np.abs(params[i].detach().numpy() - d[f"step{s}/{tag}/w/{i}"]).max() <= 2 * (s + 1) * 2e-4 * 1.01
