#!/usr/bin/env python
"""bench.py — Pix2Pix 256x256 train-step throughput (BASELINE.json metric, configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (rank 0 only)

A "step" is one Pix2Pix.train_step (reference pix2pix.py:190-218) over one synthetic batch:
generator forward, two discriminator forwards, four losses, both backward sweeps, both Keras-Adam
updates.  Workload: 256x256 RGB, GLOBAL batch 64 split evenly over the N ranks (strong scaling),
16-bit (fp16 storage with a static loss scale; GAN_B200_ACT=bf16 for bf16) tcgen05 convolutions with fp32 master
weights / statistics / optimizer; N>1: bucketed NCCL reduce-scatter + sharded Adam + all-gather (N >= 4) or
all-reduce + fused Adam (N = 2).

  value   images/s with the inputs already resident in HBM (pool of 8 distinct batches, 805 MB,
          larger than the 126 MB L2), CUDA events on the library's stream, max over ranks; the window of exactly
          K steps is repeated WINDOWS times and the MEDIAN window is reported (`windows.ms` lists all of them).
  e2e     same metric through the public API with HOST (pinned) inputs: the H2D copy of both
          images and the D2H read of the four losses are inside the timed region every step.
  roofline  dominant kernel family = tcgen05 forward/dgrad implicit-GEMM convolutions, timed live
          with CUDA events around every launch (separate profiled pass over the same steps);
          achieved = algorithmic FLOPs of the UNPADDED layers / duration; `frac` is against the measured BURST bf16
          peak (a kernel timed alone), `frac_sustained` against the sustained one; `traffic` from
          profiles/r02_traffic.json (one ncu --set full capture of the dominant kernel) when present.
  cpu_baseline  the oracle's torch-CPU fp32 restatement of the same train step on the host cores
          (TensorFlow, which the reference needs, is not installable here), bounded sample: CPU_SAMPLE_BATCH images,
          the SAME sample definition as --impl reference.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GLOBAL_BATCH = 64
SIZE, CH = 256, 3
LAMBDA = 100
SEED = 123
POOL = 8
FLOPS_PER_IMAGE = 80.546e9          # 3F_G - F_G1 + 7F_D - 2F_D1 (SURVEY 8d, 256^2 RGB)
METRIC = "pix2pix_256_train_images_per_s"


WINDOWS = 5                         # the K-step timed window is repeated; the median window is reported
CPU_SAMPLE_BATCH = 4                # ONE CPU sample definition for cpu_baseline and --impl reference


def read_peaks():
    """(burst bf16 TFLOP/s, sustained bf16 TFLOP/s, HBM GB/s, source)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return (d.get("bf16_tflops", 1679.9), d.get("bf16_tflops_sustained", 1397.7), d.get("hbm_gbs", 6547.5),
                "measured (MEASURED_PEAKS.json)")
    return 1680.0, 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def median(v):
    s = sorted(v)
    n = len(s)
    return s[n // 2] if n % 2 else 0.5 * (s[n // 2 - 1] + s[n // 2])


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons during the timed region (pynvml, else nvidia-smi)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
                     "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80}
            while not self._stop_evt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.05)
        except Exception:  # noqa: BLE001
            import subprocess
            while not self._stop_evt.is_set():
                try:
                    o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                        "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    a, b = o.strip().split(",")
                    self.samples.append(int(a)); self.max_mhz = int(b)
                except Exception:  # noqa: BLE001
                    pass
                time.sleep(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_oracle_step_rate(batch, steps, warmup, threads):
    """images/s of the oracle's torch-CPU fp32 Pix2Pix train step (the only place outside tests/
    and smoke() that executes oracle/)."""
    import numpy as np
    import torch
    from oracle import gan_oracle as O
    torch.set_num_threads(threads)
    rng_w = np.random.default_rng(SEED + 1)
    gp = O.to_torch(O.init_params(O.generator_spec(CH), rng_w, "batchnorm"), torch.float32)
    dp = O.to_torch(O.init_params(O.discriminator_spec(CH, True), rng_w, "batchnorm"), torch.float32)
    go, do = O.KerasAdam(gp), O.KerasAdam(dp)
    irng = np.random.default_rng(SEED)
    x = torch.tensor(O.synthetic_images(irng, batch, SIZE, SIZE, CH)); y = torch.tensor(O.synthetic_images(irng, batch, SIZE, SIZE, CH))
    masks = O.generator_keep_masks(SEED, 0, 0, batch, SIZE)
    times = []
    for i in range(warmup + steps):
        t = time.perf_counter()
        O.pix2pix_train_step(gp, dp, go, do, x, y, float(LAMBDA), True, masks)
        if i >= warmup:
            times.append(time.perf_counter() - t)
    total = sum(times)
    return batch * len(times) / total, 1e3 * total / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    batch = CPU_SAMPLE_BATCH
    rate, ms = cpu_oracle_step_rate(batch, args.steps, args.warmup, threads)
    sample = (f"{args.steps} train steps of batch {batch} (bounded sample of the global-batch-{GLOBAL_BATCH} workload), "
              f"torch-CPU fp32 restatement of the reference step (TensorFlow unavailable), {threads} threads")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"Pix2Pix train step 256x256x3, global batch {GLOBAL_BATCH} (CPU sample: batch {batch})",
                       "global_batch": GLOBAL_BATCH, "img_size": SIZE, "channels": CH, "lambda": LAMBDA},
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "torch_threads": torch.get_num_threads()}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from gan_b200 import Pix2Pix

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torchrun (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert GLOBAL_BATCH % world == 0
    B = GLOBAL_BATCH // world
    if args.shard_of:
        assert world == 1
        B = GLOBAL_BATCH // args.shard_of

    cfg = {"img_size": SIZE, "channels": str(CH), "learning_rate": 2e-4, "beta_1": 0.5, "beta_2": 0.999,
           "lambda": LAMBDA, "generator_loss": "l1", "seed": SEED, "precision": "bf16", "device": local,
           "batch_size": B, "epochs": 1}
    model = Pix2Pix(cfg)                       # random-init weights N(0,0.02) from default_rng(seed+1)
    ctx = model.ctx
    ctx.set_graphs(not args.no_graphs)          # whole-step CUDA graphs (eager for the profiled roofline pass)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    # synthetic inputs: U[-1,1) float32 NHWC, global sample index keyed so every world size sees the same data
    rng = np.random.default_rng(SEED)
    pool_h = []
    for _ in range(POOL):
        xg = rng.uniform(-1, 1, size=(GLOBAL_BATCH, SIZE, SIZE, CH)).astype(np.float32)
        yg = rng.uniform(-1, 1, size=(GLOBAL_BATCH, SIZE, SIZE, CH)).astype(np.float32)
        pool_h.append((torch.from_numpy(xg[rank * B:(rank + 1) * B].copy()).pin_memory(),
                       torch.from_numpy(yg[rank * B:(rank + 1) * B].copy()).pin_memory()))
    pool_d = [(x.cuda(non_blocking=False), y.cuda(non_blocking=False)) for x, y in pool_h]
    img_bytes = B * SIZE * SIZE * CH * 4

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up -------------------------------------------------------------------------------
    for i in range(max(args.warmup, 3)):
        x, y = pool_d[i % POOL]
        model.train_step(x, y, True, sync=False)
    ctx.sync()

    # ---- value: inputs resident in HBM.  The window of EXACTLY K steps (barrier + synchronize on both sides,
    #      CUDA events on the library stream, max over ranks) is repeated WINDOWS times; the median is reported.
    sampler = ClockSampler(local); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    windows_ms = []
    for w in range(WINDOWS):
        barrier()
        l0 = ctx.launch_count()
        e0.record(stream)
        for i in range(args.steps):
            x, y = pool_d[(w * args.steps + i) % POOL]
            model.train_step(x, y, True, sync=False)
        e1.record(stream)
        barrier()
        launches = ctx.launch_count() - l0
        windows_ms.append(max_over_ranks(e0.elapsed_time(e1)))
    ms_total = median(windows_ms)
    clocks = sampler.stop()
    ms_step = ms_total / args.steps
    value = GLOBAL_BATCH * args.steps / (ms_total * 1e-3)
    if args.shard_of:
        value = B * args.steps / (ms_total * 1e-3)
    last_losses = [float(v) for v in ctx.last_losses(4)]

    # ---- e2e: host (pinned) inputs, H2D + loss read-back inside the timed region ----------------
    for i in range(2):
        model.train_step(pool_h[i % POOL][0], pool_h[i % POOL][1], True)
    e2e_windows = []
    for w in range(3):
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        ctx.prefetch(*pool_h[0])
        for i in range(args.steps):
            x, y = pool_h[i % POOL]
            # tf.data-style prefetch of the NEXT batch (pix2pix.py:163): its H2D copy overlaps this step
            nxt = pool_h[(i + 1) % POOL]
            model.train_step(x, y, True, sync=False)   # consumes the prefetched device copy of (x, y)
            ctx.prefetch(*nxt)
            losses_i = ctx.last_losses(4)              # D2H read of the four losses + sync every step
        e1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e2e_windows.append(max_over_ranks(max(e0.elapsed_time(e1), wall_ms)))
    e2e_ms = median(e2e_windows)
    e2e_value = GLOBAL_BATCH * args.steps / (e2e_ms * 1e-3)

    # ---- input pipeline row: uint8 pair images from pinned host memory, split/resize/crop/
    #      flip/normalize on the device, prefetched on the copy stream while the previous step computes ----
    from gan_b200 import input_pipeline
    prng = np.random.default_rng(SEED + 3)
    pair_shape = (SIZE, 2 * SIZE, CH)
    pairs_meta = [np.empty(pair_shape, dtype=np.uint8)] * B          # only the shapes are read
    u8_pool = [torch.from_numpy(prng.integers(0, 256, size=(B, int(np.prod(pair_shape))), dtype=np.uint8)).pin_memory()
               for _ in range(POOL)]
    u8_stride = int(np.prod(pair_shape))
    jit = np.random.default_rng(SEED + 4)
    for i in range(2):
        dx, dy = model.prefetch_pairs((u8_pool[i % POOL], u8_stride), pairs_meta, True, jit)
        model.train_step(dx, dy, True, sync=False)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    nxt = model.prefetch_pairs((u8_pool[0], u8_stride), pairs_meta, True, jit)
    for i in range(args.steps):
        dx, dy = nxt
        model.train_step(dx, dy, True, sync=False)
        nxt = model.prefetch_pairs((u8_pool[(i + 1) % POOL], u8_stride), pairs_meta, True, jit)
        ctx.last_losses(4)
    e1.record(stream)
    barrier()
    u8_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
    # the gather kernel alone: CUDA events around both launches of one batch, inputs resident in HBM
    u8_dev = [t.cuda() for t in u8_pool]
    out_a = torch.empty((B, SIZE, SIZE, CH), dtype=torch.float32, device="cuda")
    out_b = torch.empty_like(out_a)
    xa, xb = model._pair_xforms(pairs_meta, True, jit)
    torch.cuda.synchronize()
    for i in range(3):
        input_pipeline.preprocess(ctx, (u8_dev[i % POOL], u8_stride), xa, CH, SIZE, out_a)
    lib_stream_evt0, lib_stream_evt1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib_stream_evt0.record(stream)
    for i in range(10):                                    # 10 batches = 20 asynchronous calls back to back
        input_pipeline.preprocess(ctx, (u8_dev[i % POOL], u8_stride), xa, CH, SIZE, out_a)
        input_pipeline.preprocess(ctx, (u8_dev[i % POOL], u8_stride), xb, CH, SIZE, out_b)
    lib_stream_evt1.record(stream)
    torch.cuda.synchronize()
    pk_ms = lib_stream_evt0.elapsed_time(lib_stream_evt1) / 10
    # algorithmic bytes: every output float written once + the source window read once (uint8)
    pipe_bytes = 2 * B * SIZE * SIZE * CH * 4 + B * u8_stride
    del u8_dev, out_a, out_b

    # ---- roofline: per-family CUDA-event timing over the same steps (separate profiled pass) ----
    peak_burst, peak_sust, peak_gbs, peak_src = read_peaks()
    peak_tf = peak_burst
    ctx.set_profile(True)
    nprof = min(args.steps, 5)
    for i in range(nprof):
        x, y = pool_d[i % POOL]
        model.train_step(x, y, True, sync=False)
    prof = ctx.profile_read()
    ctx.set_profile(False)
    fam = {k: {"ms_per_step": v[0] / nprof, "work_per_step": v[1] / nprof, "launches_per_step": v[2] / nprof}
           for k, v in prof.items() if v[2] > 0}
    uf = prof["umma_fwd"]
    roof = None
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(tpath):
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and world == 1:
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get("dram_bytes_per_launch_avg"), tj.get("source")
    if uf[2] > 0 and uf[0] > 0:
        ach = uf[1] / (uf[0] * 1e-3) / 1e12
        # every launch of the family is event-timed alone in an eager pass (SM at its boost clock between
        # launches), so the honest denominator is the BURST peak; the sustained fraction is given beside it
        roof = {"bound": "tensor", "kernel": "k_conv_fwd_umma* (tcgen05 fwd+dgrad implicit GEMM family)", "achieved": ach,
                "peak": peak_burst, "unit": "TFLOP/s", "frac": ach / peak_burst, "frac_burst": ach / peak_burst,
                "frac_sustained": ach / peak_sust, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src + ", burst (launches event-timed one by one, eager pass)",
                "algorithmic_flops": "2*M*N*K of the unpadded layers, summed per launch (engine.cu conv_flops)",
                "launches_per_step": uf[2] / nprof, "avg_launch_ms": uf[0] / uf[2],
                "flops_per_launch": uf[1] / uf[2], "share_of_step": (uf[0] / nprof) / ms_step}
    for k in ("norm", "adam", "pack"):
        if k in fam and fam[k]["ms_per_step"] > 0:
            fam[k]["achieved_GBps"] = fam[k]["work_per_step"] / (fam[k]["ms_per_step"] * 1e-3) / 1e9
            fam[k]["frac_of_hbm_peak"] = fam[k]["achieved_GBps"] / peak_gbs
    for k in ("umma_fwd", "umma_wgrad", "ffma_fwd", "ffma_wgrad"):
        if k in fam and fam[k]["ms_per_step"] > 0:
            fam[k]["achieved_TFLOPs"] = fam[k]["work_per_step"] / (fam[k]["ms_per_step"] * 1e-3) / 1e12

    # ---- CPU baseline (rank 0, N=1 only) --------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, ms = cpu_oracle_step_rate(CPU_SAMPLE_BATCH, 5, 2, threads)
        cpu = {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"5 timed + 2 warm-up train steps of batch {CPU_SAMPLE_BATCH} at 256x256x3, the same sample definition "
                         "as --impl reference (oracle torch-CPU fp32 restatement; TensorFlow, which the reference needs, "
                         "is unavailable)", "ms_per_step": ms}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "windows": {"n": WINDOWS, "ms": windows_ms, "stat": "median of WINDOWS windows of exactly `steps` steps",
                            "e2e_ms": e2e_windows},
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"Pix2Pix train step 256x256x3, global batch {GLOBAL_BATCH}, data-parallel",
                           "global_batch": GLOBAL_BATCH, "per_gpu_batch": B, "img_size": SIZE, "channels": CH,
                           "lambda": LAMBDA, "parallelism": f"dp{world}", "cuda_graphs": not args.no_graphs,
                           "l2": f"inputs larger than L2: pool of {POOL} distinct batches ({2 * POOL * img_bytes / 1e6:.0f} MB/rank)"},
                "conv_tflops_per_gpu": FLOPS_PER_IMAGE * value / world / 1e12,
                "conv_frac_of_bf16_peak": FLOPS_PER_IMAGE * value / world / 1e12 / peak_sust,
                "conv_frac_of_bf16_peak_burst": FLOPS_PER_IMAGE * value / world / 1e12 / peak_burst,
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": 2 * img_bytes * world,
                        "d2h_bytes_per_step": 16 * world, "ms_per_step": e2e_ms / args.steps},
                "input_pipeline": {"what": "uint8 pair images (256x512x3) in pinned host memory -> split, nearest resize to 286, "
                                           "random crop, mirror, normalize on the device, prefetched; then the same train step",
                                   "e2e_value": GLOBAL_BATCH * args.steps / (u8_ms * 1e-3), "unit": "images/s",
                                   "ms_per_step": u8_ms / args.steps, "h2d_bytes_per_step": B * u8_stride * world,
                                   "kernel": "k_preprocess (2 launches per batch)", "bound": "hbm",
                                   "api_ms_per_batch": pk_ms, "bytes_per_batch": pipe_bytes,
                                   "achieved_GBps": pipe_bytes / (pk_ms * 1e-3) / 1e9,
                                   "frac_of_hbm_peak": pipe_bytes / (pk_ms * 1e-3) / 1e9 / peak_gbs,
                                   "note": "CUDA events around 20 back-to-back asynchronous API calls (each = a 2.8 KB transform "
                                           "upload + one launch), so host enqueue latency is included: a lower bound on the "
                                           "kernel's bandwidth; its ncu duration is in profiles/"},
                "gpu_launches": int(launches * world),
                "roofline": roof, "families": fam, "cpu_baseline": cpu, "last_losses": last_losses}
        if args.shard_of:
            line["diagnostic"] = f"one rank's shard of a {args.shard_of}-rank job on one GPU, no collective; not a bench value"
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--shard-of", type=int, default=0,
                    help="diagnostic: on ONE GPU run the per-rank shard of a W-rank job (batch = global/W, no collective); "
                         "the line is tagged diagnostic and is not a bench value")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
