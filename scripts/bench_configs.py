"""Timing of the other BASELINE.json configurations (they are parity-test cases, not bench lines: see
tests/test_gpu_configs.py, tests/test_gpu_parity_configs.py).

    python scripts/bench_configs.py                                   # one GPU: each rank's SHARD of the 8-GPU configs
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_configs.py
                                                                      # the configs as BASELINE.json states them

configs[2]  CycleGAN 256x256x3, global batch 32 on 8 GPUs  -> 4 (x, y) pairs per rank
configs[3]  Pix2Pix 512x512 (reference default channels='1', also '3'), global batch 32 on 8 GPUs -> 4 images per rank
configs[4]  generator-only predict 256x256x3, batch 256 on ONE GPU (rank 0 only)
Every rank keeps its per-rank batch (weak definition of the shard), gradients go through the library's NCCL path
(bucketed bf16 all-reduce under the backward sweep, fused Adam; GAN_B200_SHARD_OPT=1 for the sharded optimizer); time = CUDA events on the library stream, max over ranks.  Prints one JSON
object on rank 0 with images/s, conv TFLOP/s per GPU and the fraction of the measured burst bf16 peak."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

world = int(os.environ.get("WORLD_SIZE", 1))
rank = int(os.environ.get("RANK", 0))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from gan_b200 import Pix2Pix, CycleGAN  # noqa: E402

peak = 1679.9
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk)).get("bf16_tflops", peak)


def timed(fn, stream, steps=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def images(b, s, c, seed):
    rng = np.random.default_rng(seed + 100 * rank)
    return [torch.from_numpy(rng.uniform(-1, 1, size=(b, s, s, c)).astype(np.float32)).cuda() for _ in range(2)]


def entry(ms, per_rank, gflop_per_unit, unit):
    tf = gflop_per_unit * per_rank / (ms * 1e-3) / 1e3
    return {"per_gpu_batch": per_rank, "global_batch": per_rank * world, "n_gpus": world, "ms_per_step": ms,
            f"{unit}_per_s": per_rank * world / ms * 1e3, "conv_tflops_per_gpu": tf, "frac_of_burst_bf16_peak": tf / peak}


out = {"n_gpus": world, "note": "one GPU = one rank's shard of the 8-GPU configuration" if world == 1 else "as BASELINE.json states"}
base = dict(learning_rate=2e-4, beta_1=0.5, beta_2=0.999, generator_loss='l1', seed=123, precision='bf16', epochs=1, device=local)

# configs[2]: CycleGAN 256x256x3, 4 pairs per rank
cfg = dict(base, img_size=256, channels='3', batch_size=4); cfg['lambda'] = 10
m = CycleGAN(cfg); m.ctx.set_graphs(True)
stream = torch.cuda.ExternalStream(m.ctx.stream(), device=torch.device("cuda", local))
x, y = images(4, 256, 3, 1)
ms = timed(lambda: m.train_step(x, y, True, sync=False), stream)
out["cyclegan_256_b32_on_8" if world == 8 else f"cyclegan_256_4_per_rank_x{world}"] = entry(ms, 4, 305.040, "pairs")
m.ctx.close()

# configs[3]: Pix2Pix 512x512, 4 images per rank
for ch in (1, 3):
    cfg = dict(base, img_size=512, channels=str(ch), batch_size=4); cfg['lambda'] = 100
    m = Pix2Pix(cfg); m.ctx.set_graphs(True)
    stream = torch.cuda.ExternalStream(m.ctx.stream(), device=torch.device("cuda", local))
    x, y = images(4, 512, ch, 2)
    ms = timed(lambda: m.train_step(x, y, True, sync=False), stream)
    gf = {1: 321.049, 3: 325.881}[ch]
    out[f"pix2pix_512_c{ch}_b32_on_8" if world == 8 else f"pix2pix_512_c{ch}_4_per_rank_x{world}"] = entry(ms, 4, gf, "images")
    m.ctx.close()

# configs[4]: generator-only predict, 256x256x3, batch 256 on one GPU
if rank == 0:
    saved_world, world = world, 1                                  # single-GPU measurement: no barrier / max over ranks
    cfg = dict(base, img_size=256, channels='3', batch_size=256); cfg['lambda'] = 100
    m = Pix2Pix(cfg) if saved_world == 1 else None
    if m is not None:
        stream = torch.cuda.ExternalStream(m.ctx.stream(), device=torch.device("cuda", local))
        x, _ = images(256, 256, 3, 3)
        res = torch.empty_like(x)
        ms = timed(lambda: m.generator(x, training=True, out=res), stream, steps=10, warmup=3)
        tf = 12.0964 * 256 / (ms * 1e-3) / 1e3
        out["predict_256_b256"] = {"batch": 256, "ms_per_call": ms, "images_per_s": 256 / ms * 1e3, "conv_tflops": tf,
                                   "frac_of_burst_bf16_peak": tf / peak,
                                   "note": "device-resident input and output; the call synchronises the stream"}
        m.ctx.close()
    world = saved_world
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
