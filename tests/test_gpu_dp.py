"""GPU parity of the data-parallel train step.  On 2 GPUs (skipped on single-GPU boxes): one process per GPU, NCCL
reduce-scatter of the gradient buckets + sharded Adam + all-gather inside the library; the result must equal the
oracle's data-parallel definition (per-shard step with identical weights and per-replica BatchNorm, losses and
gradients averaged — SURVEY 5.8/8e) and both ranks must end with identical weights.  On ONE GPU the same definition is
checked by running the shards one after the other (second half of this file)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
SEED = 123


def _worker(rank, world, port, q, model, precision, env=None):
    os.environ.update(env or {})               # communication-mode switches are read when the communicator is created
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gan_b200 import Pix2Pix, CycleGAN, shard_bounds
    from helpers import make_pix2pix, load_model
    from oracle import gan_oracle as O
    cfg = dict(img_size=256, channels='3', learning_rate=2e-4, beta_1=0.5, beta_2=0.999, generator_loss='l1',
               seed=SEED, precision=precision, device=rank)
    cfg['lambda'] = 100 if model == "pix2pix" else 10
    GB = 4
    rng = np.random.default_rng(SEED)
    x = O.synthetic_images(rng, GB, 256, 256, 3); y = O.synthetic_images(rng, GB, 256, 256, 3)
    lo, hi = shard_bounds(GB, rank, world)
    xt, yt = torch.tensor(x, dtype=torch.float64), torch.tensor(y, dtype=torch.float64)
    if model == "pix2pix":
        m = Pix2Pix(cfg)                      # joins the NCCL communicator through torch.distributed
        g_np, d_np = make_pix2pix(SEED + 1, 3, None)
        load_model(m.generator, g_np); load_model(m.discriminator, d_np)
        nets = [m.generator, m.discriminator]
    else:
        m = CycleGAN(cfg)
        wr = np.random.default_rng(SEED + 1)
        specs = [O.generator_spec(3), O.generator_spec(3), O.discriminator_spec(3, False), O.discriminator_spec(3, False)]
        nets_np = [O.init_params(s, wr, "instancenorm") for s in specs]
        nets = [m.generator_g, m.generator_f, m.discriminator_x, m.discriminator_y]
        for mod, arrs in zip(nets, nets_np):
            load_model(mod, arrs)
    assert m.ctx.world == world and m.ctx.rank == rank
    call0 = m.ctx.call_counter()
    w_before = [n.get_flat_params() for n in nets]
    losses = [float(v) for v in m.train_step(x[lo:hi], y[lo:hi], True)]
    res = {"rank": rank, "losses": losses, "w_sums": [float(np.abs(n.get_flat_params()).sum()) for n in nets],
           "moved": [float(np.abs(n.get_flat_params() - w0).max()) for n, w0 in zip(nets, w_before)]}
    # a second step exercises the replicas' re-packed weights (sharded optimizer: all-gather + repack)
    res["losses2"] = [float(v) for v in m.train_step(x[lo:hi], y[lo:hi], True)]
    res["w_sums2"] = [float(np.abs(n.get_flat_params()).sum()) for n in nets]
    if rank == 0:
        # oracle: data-parallel definition with the same global-sample-keyed dropout masks
        if model == "pix2pix":
            gp, dp = O.to_torch(g_np, torch.float64), O.to_torch(d_np, torch.float64)
            go, do = O.KerasAdam(gp), O.KerasAdam(dp)
            masks = O.generator_keep_masks(SEED, call0, 0, GB, 256)
            ref_losses, ref_gg, ref_dg = O.pix2pix_train_step(gp, dp, go, do, xt, yt, 100.0, True, masks, world=world)
            masks2 = O.generator_keep_masks(SEED, call0 + 1, 0, GB, 256)
            ref_losses2, _, _ = O.pix2pix_train_step(gp, dp, go, do, xt, yt, 100.0, True, masks2, world=world)
            ref_w = [np.concatenate([p.detach().numpy().ravel() for p in gp]), np.concatenate([p.detach().numpy().ravel() for p in dp])]
        else:
            calls = ['fake_y', 'cycled_x', 'fake_x', 'cycled_y', 'same_x', 'same_y']
            tn = [O.to_torch(a, torch.float64) for a in nets_np]
            opts = [O.KerasAdam(p) for p in tn]
            masks = {n: O.generator_keep_masks(SEED, call0 + i, 0, GB, 256) for i, n in enumerate(calls)}
            ref_losses, _ = O.cyclegan_train_step(tn, opts, xt, yt, 10.0, True, masks, world=1)     # InstanceNorm: DP == unsharded
            masks2 = {n: O.generator_keep_masks(SEED, call0 + 6 + i, 0, GB, 256) for i, n in enumerate(calls)}
            ref_losses2, _ = O.cyclegan_train_step(tn, opts, xt, yt, 10.0, True, masks2, world=1)
            ref_w = [np.concatenate([p.detach().numpy().ravel() for p in t]) for t in tn]
        res["ref_losses"] = ref_losses; res["ref_losses2"] = ref_losses2
        # after two Keras-Adam steps every weight sits within 2*2*lr of the oracle's (sign-like first updates)
        res["w_err"] = [float(np.abs(n.get_flat_params() - r).max()) for n, r in zip(nets, ref_w)]
    q.put(res)
    dist.barrier()
    m.ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("model,precision,env", [
    ("pix2pix", "fp32", {}),                                  # fp32 parity mode: fp32 gradient all-reduce + fused Adam
    ("pix2pix", "bf16", {}),                                  # 16-bit mode (default): gradient buckets travel as bf16
    ("pix2pix", "bf16", {"GAN_B200_COMM16": "0"}),            # 16-bit mode, fp32 on the wire
    ("pix2pix", "bf16", {"GAN_B200_SHARD_OPT": "1"}),         # reduce-scatter + 1/world Adam + all-gather (opt-in)
    ("cyclegan", "fp32", {}),
    ("cyclegan", "fp32", {"GAN_B200_SHARD_OPT": "1"}),
])
def test_two_gpu_data_parallel_step_matches_oracle(model, precision, env):
    """One process per GPU, NCCL all-reduce of the gradient buckets under the backward sweep (or the opt-in sharded
    optimizer): both ranks report the same (all-reduced) losses, end with identical weights, and match the oracle's
    data-parallel definition for two consecutive steps."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, model, precision, env)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=900) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    out.sort(key=lambda r: r["rank"])
    r0, r1 = out
    assert r0["losses"] == r1["losses"] and r0["losses2"] == r1["losses2"]        # losses are all-reduced means
    assert r0["w_sums"] == r1["w_sums"] and r0["w_sums2"] == r1["w_sums2"]        # replicas stay identical
    assert all(mv > 0 for mv in r0["moved"])                                       # every net was updated
    tol = 1e-4 if precision == "fp32" else 1e-2
    for a, r in zip(r0["losses"] + r0["losses2"], r0["ref_losses"] + r0["ref_losses2"]):
        assert abs(a - r) <= tol * max(1.0, abs(r)), (r0["losses"], r0["ref_losses"], r0["losses2"], r0["ref_losses2"])
    # two Keras-Adam steps: |update| <= lr * (1 + small) each, for the device and for the oracle
    assert all(e <= 2 * 2 * 2e-4 * (1.01 if precision == "fp32" else 1.05) for e in r0["w_err"]), r0["w_err"]


# ---------------------------------------------------------------------------------------------------------------
# The same data-parallel DEFINITION on ONE GPU (so that a single-GPU box still checks it): the two shards of a global
# batch run one after the other in two contexts with identical weights, each with its global sample offset
# (gan_ctx_set_sample_offset keys the dropout masks on the GLOBAL sample index); gradients and losses are averaged on
# the host as the all-reduce + 1/world scaling would.
# ---------------------------------------------------------------------------------------------------------------
def _shard_models(cls, cfg, loaders, world):
    ms = []
    for r in range(world):
        m = cls(dict(cfg))
        loaders(m)
        ms.append(m)
    return ms


def test_one_gpu_sharded_pix2pix_equals_data_parallel_oracle():
    """Pix2Pix: per-replica BatchNorm statistics, global-sample-keyed dropout."""
    from gan_b200 import Pix2Pix, shard_bounds
    from helpers import make_pix2pix, load_model
    from oracle import gan_oracle as O
    world, GB = 2, 4
    cfg = dict(img_size=256, channels='3', learning_rate=2e-4, beta_1=0.5, beta_2=0.999, generator_loss='l1',
               seed=SEED, precision='fp32', device=0)
    cfg['lambda'] = 100
    g_np, d_np = make_pix2pix(SEED + 1, 3, None)
    ms = _shard_models(Pix2Pix, cfg, lambda m: (load_model(m.generator, g_np), load_model(m.discriminator, d_np)), world)
    rng = np.random.default_rng(SEED)
    x = O.synthetic_images(rng, GB, 256, 256, 3); y = O.synthetic_images(rng, GB, 256, 256, 3)
    losses, gg, dg = [], [], []
    for r, m in enumerate(ms):
        lo, hi = shard_bounds(GB, r, world)
        m.ctx.set_sample_offset(lo)
        assert m.ctx.call_counter() == 0
        losses.append([float(v) for v in m.train_step(x[lo:hi], y[lo:hi], True)])
        gg.append(m.generator.get_flat_grads().astype(np.float64)); dg.append(m.discriminator.get_flat_grads().astype(np.float64))
    dev_losses = np.mean(np.array(losses), axis=0)
    dev_gg, dev_dg = sum(gg) / world, sum(dg) / world
    gp, dp = O.to_torch(g_np, torch.float64), O.to_torch(d_np, torch.float64)
    go, do = O.KerasAdam(gp), O.KerasAdam(dp)
    masks = O.generator_keep_masks(SEED, 0, 0, GB, 256)
    ref_losses, ref_gg, ref_dg = O.pix2pix_train_step(gp, dp, go, do, torch.tensor(x, dtype=torch.float64),
                                                      torch.tensor(y, dtype=torch.float64), 100.0, True, masks, world=world)
    for a, r in zip(dev_losses, ref_losses):
        assert abs(a - r) <= 1e-4 * max(1.0, abs(r)), (dev_losses, ref_losses)
    for dev, ref in ((dev_gg, ref_gg), (dev_dg, ref_dg)):
        ref_flat = np.concatenate([g.numpy().ravel() for g in ref])
        assert np.abs(dev - ref_flat).max() / np.abs(ref_flat).max() < 3e-3     # batch-2 shards: fp32 conditioning level
    # and the UNSHARDED batch-4 step is a different function (BatchNorm couples the batch): the test would notice
    # a kernel that silently normalised over the wrong set
    ref1, _, _ = O.pix2pix_train_step(O.to_torch(g_np, torch.float64), O.to_torch(d_np, torch.float64), O.KerasAdam(gp), O.KerasAdam(dp),
                                      torch.tensor(x, dtype=torch.float64), torch.tensor(y, dtype=torch.float64), 100.0, False, masks, world=1)
    assert max(abs(a - b) for a, b in zip(ref1, ref_losses)) > 1e-3
    for m in ms:
        m.ctx.close()


def test_one_gpu_sharded_cyclegan_equals_unsharded_oracle():
    """CycleGAN: InstanceNorm is per sample, so the sharded step equals the UNSHARDED oracle batch (bf16-free fp32 mode)."""
    from gan_b200 import CycleGAN, shard_bounds
    from helpers import load_model
    from oracle import gan_oracle as O
    world, GB = 2, 2
    cfg = dict(img_size=256, channels='3', learning_rate=2e-4, beta_1=0.5, beta_2=0.999, seed=SEED, precision='fp32', device=0)
    cfg['lambda'] = 10
    rng = np.random.default_rng(SEED + 1)
    specs = [O.generator_spec(3), O.generator_spec(3), O.discriminator_spec(3, False), O.discriminator_spec(3, False)]
    nets_np = [O.init_params(s, rng, "instancenorm") for s in specs]

    def load(m):
        for mod, arrs in zip([m.generator_g, m.generator_f, m.discriminator_x, m.discriminator_y], nets_np):
            load_model(mod, arrs)
    ms = _shard_models(CycleGAN, cfg, load, world)
    irng = np.random.default_rng(SEED)
    x = O.synthetic_images(irng, GB, 256, 256, 3); y = O.synthetic_images(irng, GB, 256, 256, 3)
    losses, grads = [], []
    for r, m in enumerate(ms):
        lo, hi = shard_bounds(GB, r, world)
        m.ctx.set_sample_offset(lo)
        losses.append([float(v) for v in m.train_step(x[lo:hi], y[lo:hi], True)])
        grads.append([mod.get_flat_grads().astype(np.float64) for mod in (m.generator_g, m.generator_f, m.discriminator_x, m.discriminator_y)])
    dev_losses = np.mean(np.array(losses), axis=0)
    calls = ['fake_y', 'cycled_x', 'fake_x', 'cycled_y', 'same_x', 'same_y']
    masks = {n: O.generator_keep_masks(SEED, i, 0, GB, 256) for i, n in enumerate(calls)}
    nets = [O.to_torch(a, torch.float64) for a in nets_np]
    ref_losses, g1, g2, g3, g4, _ = O.cyclegan_losses_and_grads(nets[0], nets[1], nets[2], nets[3], torch.tensor(x, dtype=torch.float64),
                                                                 torch.tensor(y, dtype=torch.float64), 10.0, masks)
    for a, r in zip(dev_losses, ref_losses):
        assert abs(a - float(r)) <= 1e-4 * max(1.0, abs(float(r))), (dev_losses, [float(v) for v in ref_losses])
    for k, ref in enumerate((g1, g2, g3, g4)):
        dev = sum(g[k] for g in grads) / world
        ref_flat = np.concatenate([g.numpy().ravel() for g in ref])
        assert np.abs(dev - ref_flat).max() / np.abs(ref_flat).max() < 3e-3, k
    for m in ms:
        m.ctx.close()
