"""CPU tests of the host-side logic, including the world_size-2 data-parallel path over gloo."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_and_loss_value():
    from gan_b200 import shard_bounds, LossValue
    assert [shard_bounds(64, r, 8) for r in (0, 7)] == [(0, 8), (56, 64)]
    with pytest.raises(ValueError):
        shard_bounds(10, 0, 4)
    v = LossValue(1.5)
    assert v.numpy().tolist() == 1.5 and isinstance(v.numpy(), np.float32)


def test_loss_dict_keys_match_reference():
    from gan_b200.utils import pix2pix_losses, cyclegan_losses
    assert list(pix2pix_losses()) == ['Generator Total Loss', 'Generator Loss (Primary)',
                                      'Generator Loss (Secondary)', 'Discriminator Loss']
    assert len(cyclegan_losses()) == 7 and 'Total Cycle Loss' in cyclegan_losses()


def test_host_bce_matches_oracle():
    from gan_b200.base_gan import BinaryCrossentropyFromLogits
    from oracle import gan_oracle as O
    x = np.random.default_rng(0).normal(size=(2, 30, 30, 1)) * 4
    f = BinaryCrossentropyFromLogits()
    assert abs(f(np.ones_like(x), x) - float(O.bce_from_logits(torch.tensor(x), 1.0))) < 1e-12
    assert abs(f(np.zeros_like(x), x) - float(O.bce_from_logits(torch.tensor(x), 0.0))) < 1e-12


def _dp_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from gan_b200 import exchange_unique_id, shard_bounds
    from oracle import gan_oracle as O
    # 1. unique-id plumbing: rank 0's id reaches every rank
    r, w, uid = exchange_unique_id(make_id=lambda: bytes(range(128)))
    assert (r, w) == (rank, world) and uid == bytes(range(128))
    # 2. data-parallel definition: per-shard reference step with identical weights,
    #    gradients and losses averaged with an all-reduce == oracle's world=2 restatement.
    rng = np.random.default_rng(5)
    d_np = O.init_params(O.discriminator_spec(1, True), rng, "batchnorm")
    B = 4
    x = O.synthetic_images(rng, B, 32, 32, 1); y = O.synthetic_images(rng, B, 32, 32, 1)
    lo, hi = shard_bounds(B, rank, world)
    dp = O.to_torch(d_np)
    logits = O.discriminator_forward(dp, torch.tensor(x[lo:hi], dtype=torch.float64), torch.tensor(y[lo:hi], dtype=torch.float64))
    loss = O.bce_from_logits(logits, 1.0)
    grads = torch.autograd.grad(loss, dp)
    flat = torch.cat([g.flatten() for g in grads] + [loss.detach().reshape(1)])
    dist.all_reduce(flat)            # the exchange step of the path: sum over ranks ...
    flat /= world                    # ... divided by world before Adam
    if rank == 0:
        ref_flat = []
        dp2 = O.to_torch(d_np)
        acc = None
        for s in range(world):
            l2 = O.bce_from_logits(O.discriminator_forward(dp2, torch.tensor(x[s * 2:(s + 1) * 2], dtype=torch.float64),
                                                           torch.tensor(y[s * 2:(s + 1) * 2], dtype=torch.float64)), 1.0)
            g2 = torch.autograd.grad(l2, dp2)
            v = torch.cat([g.flatten() for g in g2] + [l2.detach().reshape(1)])
            acc = v if acc is None else acc + v
        q.put(float((flat - acc / world).abs().max()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_host_path_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=10) < 1e-12


def test_checkpoint_manager_numbering_and_rotation(tmp_path):
    """CheckpointManager (role of tf.train.CheckpointManager, pix2pix.py:418): numbered files, newest
    wins, old ones rotate out — host logic only, the tensor payload is stubbed."""
    import numpy as np
    from gan_b200 import checkpoint as ck

    class Stub(ck.Checkpoint):
        def __init__(self):
            self._objects, self.ctx = {}, None

        def _collect(self):
            return {"x": np.arange(3)}
    mgr = ck.CheckpointManager(Stub(), str(tmp_path), max_to_keep=1)
    assert mgr.latest_checkpoint is None
    p1 = mgr.save(); p2 = mgr.save()
    assert p2.endswith("ckpt-2.npz") and ck.latest_checkpoint(str(tmp_path)) == p2
    import os
    assert not os.path.exists(p1) and [os.path.basename(f) for f in mgr.checkpoints] == ["ckpt-2.npz"]
    (tmp_path / "ckpt-10.npz").write_bytes(open(p2, "rb").read())
    assert ck.latest_checkpoint(str(tmp_path)).endswith("ckpt-10.npz")          # numeric, not lexicographic, order
    assert not any(f.endswith(".tmp.npz") for f in os.listdir(tmp_path))


def test_input_pipeline_host_side_descriptions():
    """Host half of the device input pipeline (gan_b200/input_pipeline.py): the random draws of
    random_jitter (pix2pix.py:66-90) stay in range, transforms describe train vs prediction paths the way
    the reference applies them, ragged images pack into one strided buffer.  No device work."""
    import numpy as np
    from gan_b200 import input_pipeline as ip

    rng = np.random.default_rng(0)
    draws = [ip.draw_jitter(rng, 256) for _ in range(2000)]
    assert all(0 <= cy <= 30 and 0 <= cx <= 30 for cy, cx, _ in draws)
    assert {cy for cy, _, _ in draws} == set(range(31)) and {cx for _, cx, _ in draws} == set(range(31))
    flips = sum(f for _, _, f in draws)
    assert 850 < flips < 1150                                   # mirror with probability 1/2

    t = ip.xform(300, 500, 250, 250, 256, True, (7, 9), True)
    assert (t.src_h, t.src_w, t.col0, t.cols, t.pre, t.mid, t.crop_y, t.crop_x, t.flip) == (300, 500, 250, 250, 0, 286, 7, 9, 1)
    p = ip.xform(300, 500, 0, 250, 256, False, (7, 9), True)    # prediction path ignores the draws
    assert (p.mid, p.crop_y, p.crop_x, p.flip) == (0, 0, 0, 0)
    c = ip.xform(199, 301, 0, 301, 256, True, (0, 30), False, pre=256)   # CycleGAN load(resize=True)
    assert (c.pre, c.mid, c.crop_x) == (256, 286, 30)

    ims = [np.full((4, 6, 3), 1, np.uint8), np.full((5, 8, 3), 2, np.uint8)]
    buf, stride = ip.pack_images(ims)
    assert stride == 120 and buf.shape == (2, 120) and buf[0, :72].min() == 1 and buf[0, 72:].max() == 0 and buf[1].min() == 2
    import pytest
    with pytest.raises(ValueError):
        ip.pack_images([np.zeros((4, 4, 3), np.float32)])


def test_profile_scripts_parse_the_committed_launch_list(tmp_path):
    """scripts/summarize_launches.py and scripts/make_traffic.py (the tools behind profiles/*.md and roofline.traffic) read
    the committed ncu launch list: 200+ launches, the forward-type convolution family present, positive DRAM traffic."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csv_path = os.path.join(root, "profiles", "r02_launches_one_step.csv")
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "summarize_launches.py"), csv_path], capture_output=True,
                         text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    first = out.stdout.splitlines()[0]
    assert int(first.split()[0]) >= 200 and "k_conv_fwd_umma2<256>" in out.stdout and "k_conv_first_fwd" in out.stdout
    with open(os.path.join(root, "profiles", "r02_traffic.json")) as f:
        tj = json.load(f)
    assert tj["launches_captured"] >= 50 and tj["dram_bytes_per_launch_avg"] > 1e7 and tj["wgrad_launches_captured"] >= 20
