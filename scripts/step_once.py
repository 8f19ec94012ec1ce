"""Runs a few EAGER train steps for an ncu launch list: python scripts/step_once.py [steps] [batch] [pix2pix|cyclegan].
Default: Pix2Pix at the benchmarked configuration (256x256x3, batch 64).  Prints the launch count of every step."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gan_b200 import Pix2Pix, CycleGAN  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
kind = sys.argv[3] if len(sys.argv) > 3 else "pix2pix"
cfg = {"img_size": 256, "channels": "3", "learning_rate": 2e-4, "beta_1": 0.5, "beta_2": 0.999,
       "lambda": 100 if kind == "pix2pix" else 10, "generator_loss": "l1", "seed": 123, "precision": "bf16", "device": 0}
m = Pix2Pix(cfg) if kind == "pix2pix" else CycleGAN(cfg)
rng = np.random.default_rng(0)
x = rng.uniform(-1, 1, size=(B, 256, 256, 3)).astype(np.float32)
y = rng.uniform(-1, 1, size=(B, 256, 256, 3)).astype(np.float32)
import torch  # noqa: E402  (device buffers only)
xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
for i in range(steps):
    l0 = m.ctx.launch_count()
    m.train_step(xd, yd, True)
    print("step", i, "launches", m.ctx.launch_count() - l0, flush=True)
m.ctx.close()
