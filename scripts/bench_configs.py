"""Timing of the other BASELINE.json configurations on ONE GPU (they are parity-test cases, not bench
lines: see tests/test_gpu_configs.py).  Prints one JSON object; per-rank shards where the configuration
names 8 GPUs.  Usage: python scripts/bench_configs.py > profiles/rNN_other_configs.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gan_b200 import Pix2Pix, CycleGAN  # noqa: E402


def timed(fn, steps=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def images(b, s, c, seed):
    rng = np.random.default_rng(seed)
    return [torch.from_numpy(rng.uniform(-1, 1, size=(b, s, s, c)).astype(np.float32)).cuda() for _ in range(2)]


out = {}
base = dict(learning_rate=2e-4, beta_1=0.5, beta_2=0.999, generator_loss='l1', seed=123, precision='bf16', epochs=1)

# configs[2]: CycleGAN 256x256x3, global batch 32 on 8 GPUs -> 4 pairs per rank
cfg = dict(base, img_size=256, channels='3', batch_size=4); cfg['lambda'] = 10
m = CycleGAN(cfg); m.ctx.set_graphs(True)
stream = torch.cuda.ExternalStream(m.ctx.stream(), device=torch.device("cuda", 0))
x, y = images(4, 256, 3, 1)
ms = timed(lambda: m.train_step(x, y, True, sync=False))
out["cyclegan_256_shard_of_8"] = {"per_gpu_batch": 4, "ms_per_step": ms, "pairs_per_s_per_gpu": 4 / ms * 1e3,
                                  "conv_tflops": 305.040e9 * 4 / (ms * 1e-3) / 1e12}
m.ctx.close()

# configs[3]: Pix2Pix 512x512 (reference default channels='1'), global batch 32 on 8 GPUs -> 4 per rank
for ch in (1, 3):
    cfg = dict(base, img_size=512, channels=str(ch), batch_size=4); cfg['lambda'] = 100
    m = Pix2Pix(cfg); m.ctx.set_graphs(True)
    stream = torch.cuda.ExternalStream(m.ctx.stream(), device=torch.device("cuda", 0))
    x, y = images(4, 512, ch, 2)
    ms = timed(lambda: m.train_step(x, y, True, sync=False))
    gf = {1: 321.049e9, 3: 325.881e9}[ch]
    out[f"pix2pix_512_c{ch}_shard_of_8"] = {"per_gpu_batch": 4, "ms_per_step": ms, "images_per_s_per_gpu": 4 / ms * 1e3,
                                           "conv_tflops": gf * 4 / (ms * 1e-3) / 1e12}
    m.ctx.close()

# configs[4]: generator-only predict, 256x256x3, batch 256 on one GPU
cfg = dict(base, img_size=256, channels='3', batch_size=256); cfg['lambda'] = 100
m = Pix2Pix(cfg)
stream = torch.cuda.ExternalStream(m.ctx.stream(), device=torch.device("cuda", 0))
x, _ = images(256, 256, 3, 3)
res = torch.empty_like(x)
ms = timed(lambda: m.generator(x, training=True, out=res), steps=10, warmup=3)
out["predict_256_b256"] = {"batch": 256, "ms_per_call": ms, "images_per_s": 256 / ms * 1e3,
                           "conv_tflops": 12.0964e9 * 256 / (ms * 1e-3) / 1e12,
                           "note": "device-resident input and output; the call synchronises the stream"}
m.ctx.close()
print(json.dumps(out))
