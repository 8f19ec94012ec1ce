"""Golden fixture for the BENCHMARKED configuration (BASELINE.json configs[1] at N=1: Pix2Pix 256x256 RGB,
batch 64) and the calibrator of the bf16 tolerance.

    python tests/golden/make_golden_b64.py            # ~10 min on 8 host cores, writes pix2pix_b64_cal.npz

The CPU oracle needs minutes per float64 step at batch 64, so it runs HERE once and the GPU test
(tests/test_gpu_parity_configs.py) compares against this file.  Protocol (mirrored by the test, dropout call
counter in brackets): out0 = G(x) [0]; N = 10 train steps [1..10]; outN = G(x) [11].  Stored:

  * float64 oracle: out0 / outN for SAMPLES of the batch (BatchNorm couples the whole batch, the
    comparison is per stored sample), the four losses of every step, a fixed random subsample of every
    gradient tensor at step 1 (cosines), per-tensor gradient max-abs;
  * float32 oracle, free-running from the same weights, inputs and masks: its losses per step and its
    deviation from the float64 trajectory (max-rel and L2 on outN) — the CALIBRATOR: whatever torch float32
    cannot hold after N steps is not a property of the device's bf16 path.

PARITY UNPINNED against TensorFlow (oracle/gan_oracle.py header): this pins device == oracle, not oracle == TF.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import gan_oracle as O  # noqa: E402

SEED, B, SIZE, CH, N = 123, 64, 256, 3, 10
SAMPLES = (0, 21, 63)
SUB = 4096


def run(dtype):
    rng = np.random.default_rng(SEED + 1)
    g_np = O.init_params(O.generator_spec(CH), rng, "batchnorm")
    d_np = O.init_params(O.discriminator_spec(CH, True), rng, "batchnorm")
    gp, dp = O.to_torch(g_np, dtype), O.to_torch(d_np, dtype)
    go, do = O.KerasAdam(gp), O.KerasAdam(dp)
    irng = np.random.default_rng(SEED)
    x = torch.tensor(O.synthetic_images(irng, B, SIZE, SIZE, CH), dtype=dtype)
    y = torch.tensor(O.synthetic_images(irng, B, SIZE, SIZE, CH), dtype=dtype)
    res = {}
    with torch.no_grad():
        res["out0"] = O.generator_forward(gp, x, "batchnorm", O.generator_keep_masks(SEED, 0, 0, B, SIZE)).numpy().astype(np.float64)
    losses = []
    for s in range(N):
        t = time.time()
        masks = O.generator_keep_masks(SEED, 1 + s, 0, B, SIZE)
        l, gg, dg = O.pix2pix_train_step(gp, dp, go, do, x, y, 100.0, True, masks)
        losses.append([float(v) for v in l])
        if s == 0:
            res["gg"] = [g.detach().numpy().astype(np.float64) for g in gg]
            res["dg"] = [g.detach().numpy().astype(np.float64) for g in dg]
        print(dtype, "step", s, losses[-1], f"{time.time() - t:.1f}s", flush=True)
    res["losses"] = np.array(losses)
    with torch.no_grad():
        res["outN"] = O.generator_forward(gp, x, "batchnorm", O.generator_keep_masks(SEED, 1 + N, 0, B, SIZE)).numpy().astype(np.float64)
    return res


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    r64 = run(torch.float64)
    r32 = run(torch.float32)
    out = {"protocol": np.array([SEED, B, SIZE, CH, N]), "samples": np.array(SAMPLES),
           "losses64": r64["losses"], "losses32": r32["losses"]}
    for k in ("out0", "outN"):
        out[k + "_64"] = r64[k][list(SAMPLES)].astype(np.float32)
        d = r32[k] - r64[k]
        out[k + "_f32_vs_f64"] = np.array([np.abs(d).max() / np.abs(r64[k]).max(),
                                           np.linalg.norm(d) / np.linalg.norm(r64[k])])
        out[k + "_norms"] = np.array([np.abs(r64[k]).max(), np.linalg.norm(r64[k]),
                                      np.linalg.norm(r64[k][list(SAMPLES)])])
    prng = np.random.default_rng(SEED + 7)
    for tag in ("gg", "dg"):
        idx, val, mx, cos32 = [], [], [], []
        for g64, g32 in zip(r64[tag], r32[tag]):
            flat = g64.reshape(-1)
            ii = np.sort(prng.choice(flat.size, size=min(SUB, flat.size), replace=False))
            idx.append(ii.astype(np.int64)); val.append(flat[ii].astype(np.float32)); mx.append(np.abs(flat).max())
            a, b = flat, g32.reshape(-1)
            cos32.append(float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300)))
        out[tag + "_idx"] = np.concatenate(idx); out[tag + "_val"] = np.concatenate(val)
        out[tag + "_len"] = np.array([len(i) for i in idx]); out[tag + "_maxabs"] = np.array(mx)
        out[tag + "_cos_f32_vs_f64"] = np.array(cos32)
    path = os.path.join(HERE, "pix2pix_b64_cal.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")
    print("float32 oracle vs float64 oracle: out0", out["out0_f32_vs_f64"], "outN", out["outN_f32_vs_f64"])


if __name__ == "__main__":
    main()
