"""The callers either side of the step: image_pipeline on a directory of PNG pair images
(PIL decode on the host, pixel work on the device), the fit loop with the reference's checkpoint cadence
(pix2pix.py:248-323), predict."""
import os

import numpy as np
import pytest

from oracle import pipeline_oracle as P

pytestmark = pytest.mark.gpu


def _write_pairs(path, n, rng):
    from PIL import Image
    imgs = {}
    for i in range(n):
        a = rng.integers(0, 256, size=(64 + 8 * i, 2 * (80 + 4 * i), 3), dtype=np.uint8)
        name = f"pair_{i:02d}.png"
        Image.fromarray(a).save(os.path.join(path, name))
        imgs[name] = a
    return imgs


def test_image_pipeline_fit_checkpoint_predict(tmp_path):
    from gan_b200 import Pix2Pix, Checkpoint, CheckpointManager, latest_checkpoint
    rng = np.random.default_rng(0)
    data = tmp_path / "data"; data.mkdir()
    imgs = _write_pairs(str(data), 6, rng)
    cfg = dict(img_size=256, channels='3', learning_rate=2e-4, beta_1=0.5, beta_2=0.999, generator_loss='l1', seed=123,
               precision='bf16', epochs=6, batch_size=2, data=str(data), test_img=1, validation_size=0.2,
               input_img_orient='left')
    cfg['lambda'] = 100
    m = Pix2Pix(cfg)
    train, val, test = m.image_pipeline(predict=False)
    assert len(train) == 2 and len(val) == 1 and len(test) == 1            # 6 files: 1 test, ceil(5*0.2)=1 val, 4 train
    # validation batches are the deterministic prediction path: bit-exact against the per-image oracle
    (vx, vy), = list(val)
    import random
    random.seed(cfg['seed'])
    names = sorted(imgs) if False else [i for i in os.listdir(str(data)) if 'png' in i or 'jpg' in i]
    t = random.sample(names, 1)
    v = random.sample([i for i in names if i not in t], 1)
    ra, rb = P.pix2pix_process_pred(imgs[v[0]], 'left', 256)
    assert np.array_equal(vx[0], ra) and np.array_equal(vy[0], rb)
    for x, y in train:
        assert x.shape == (2, 256, 256, 3) and x.dtype == np.float32 and -1.0 <= x.min() and x.max() <= 1.0

    ckdir = tmp_path / "training_checkpoints"
    ck = Checkpoint(generator_optimizer=m.generator_optimizer, discriminator_optimizer=m.discriminator_optimizer,
                    generator=m.generator, discriminator=m.discriminator)
    mgr = CheckpointManager(ck, str(ckdir), max_to_keep=1)
    tr, va = m.fit(train, val, test, output_path=str(tmp_path), checkpoint_manager=mgr)
    keys = ["Generator Total Loss", "Generator Loss (Primary)", "Generator Loss (Secondary)", "Discriminator Loss"]
    assert list(tr.keys()) == keys and all(len(tr[k]) == 6 and len(va[k]) == 6 for k in keys)
    assert all(np.isfinite(vv) for k in keys for vv in tr[k] + va[k])
    assert m.generator_optimizer.iterations == 6 * 2
    assert os.path.basename(latest_checkpoint(str(ckdir))) == "ckpt-2.npz"     # saved at epoch 5 and at the last epoch
    # sample image of the first test example every 5 epochs except the last (pix2pix.py:307-313)
    from PIL import Image
    png = os.path.join(str(tmp_path), "test_images", "epoch_5.png")
    assert os.path.exists(png) and Image.open(png).size == (3 * 256 + 2 * 8, 256)
    assert not os.path.exists(os.path.join(str(tmp_path), "test_images", "epoch_6.png"))
    # metrics dump of the run driver (pix2pix.py:436-440): same keys, one mean per epoch
    import json
    from gan_b200.utils import dump_metrics
    ptr, pva = dump_metrics(os.path.join(str(tmp_path), "logs"), tr, va)
    assert os.path.basename(ptr) == "train_metrics.json" and os.path.basename(pva) == "val_metrics.json"
    with open(ptr) as f:
        back = json.load(f)
    assert list(back.keys()) == keys and back[keys[0]] == tr[keys[0]]

    pred_ds, _, _ = m.image_pipeline(predict=True)
    outs = m.predict(pred_ds, output_path=str(tmp_path))
    assert len(outs) == 6 and outs[0].shape == (1, 256, 256, 3) and np.abs(outs[0]).max() <= 1.0
    assert os.path.exists(os.path.join(str(tmp_path), "prediction_images", "img5.png"))
    m.ctx.close()


def test_cyclegan_image_pipeline_split_and_batches(tmp_path):
    """CycleGAN.image_pipeline (cycle_gan.py:87-152): unpaired directories, seeded split, batches without
    drop_remainder; the test split is the deterministic prediction path -> bit-exact against the per-image oracle."""
    import random
    from PIL import Image
    from gan_b200 import CycleGAN
    rng = np.random.default_rng(1)
    dx, dy = tmp_path / "X", tmp_path / "Y"; dx.mkdir(); dy.mkdir()
    imgs = {}
    for d, n in ((dx, 6), (dy, 5)):
        for i in range(n):
            a = rng.integers(0, 256, size=(70 + 6 * i, 90 + 5 * i, 3), dtype=np.uint8)
            Image.fromarray(a).save(os.path.join(str(d), f"im_{i}.png")); imgs[(str(d), f"im_{i}.png")] = a
    cfg = dict(img_size=256, channels='3', seed=7, precision='bf16', epochs=1, batch_size=2, input_images=str(dx),
               target_images=str(dy), test_img=1, validation_size=0.25)
    cfg['lambda'] = 10
    m = CycleGAN(cfg)
    train_X, train_Y, val_X, val_Y, test = m.image_pipeline(False)
    # 6 X files: 1 test, ceil(5*.25)=2 val, 3 train; 5 Y files: ceil(5*.25)=2 val, 3 train
    assert (len(train_X), len(train_Y), len(val_X), len(val_Y), len(test)) == (2, 2, 1, 1, 1)
    names_x = [i for i in os.listdir(str(dx)) if 'png' in i or 'jpg' in i]
    random.seed(cfg['seed'])
    t = random.sample(names_x, 1)
    (tb,) = list(test)
    assert tb.shape == (1, 256, 256, 3) and np.array_equal(tb[0], P.cyclegan_process_pred(imgs[(str(dx), t[0])], 256))
    sizes = [b.shape[0] for b in train_X]
    assert sizes == [2, 1]                                         # ragged tail batch, no drop_remainder
    for b in train_Y:
        assert b.dtype == np.float32 and -1.0 <= b.min() and b.max() <= 1.0
    pred, *rest = m.image_pipeline(True)
    assert all(r is None for r in rest) and len(list(pred)) == 6
    m.ctx.close()


def test_ssim_generator_loss_option_as_the_reference_wrote_it():
    """generator_loss='ssim' (pix2pix.py:182-186): gan_loss2 = tf.image.ssim(input, target) is a per-image vector that
    does not depend on the generator; total = gan_loss + lambda*ssim is a vector too, and tape.gradient differentiates
    its SUM over the batch -> generator gradient = batch * d(gan_loss), no L1 term.  Checked against torch autograd."""
    import torch
    from gan_b200 import Pix2Pix
    from helpers import make_pix2pix, load_model
    from oracle import gan_oracle as O
    B = 2
    cfg = dict(img_size=256, channels='3', learning_rate=2e-4, beta_1=0.5, beta_2=0.999, generator_loss='ssim', seed=123,
               precision='fp32')
    cfg['lambda'] = 100
    m = Pix2Pix(cfg)
    g_np, d_np = make_pix2pix(124, 3, None)
    load_model(m.generator, g_np); load_model(m.discriminator, d_np)
    irng = np.random.default_rng(3)
    x = O.synthetic_images(irng, B, 256, 256, 3); y = O.synthetic_images(irng, B, 256, 256, 3)
    masks = O.generator_keep_masks(123, m.ctx.call_counter(), 0, B, 256)
    total, gan, sec, disc = m.train_step(x, y, True)
    assert np.asarray(total).shape == (B,) and np.asarray(sec).shape == (B,)          # vectors, as in the reference
    gp, dp = O.to_torch(g_np, torch.float64), O.to_torch(d_np, torch.float64)
    xt, yt = torch.tensor(x, dtype=torch.float64), torch.tensor(y, dtype=torch.float64)
    gen_out = O.generator_forward(gp, xt, "batchnorm", masks)
    fake = O.discriminator_forward(dp, xt, gen_out, "batchnorm")
    gan_ref = O.bce_from_logits(fake, 1.0)
    ref_g = torch.autograd.grad(B * gan_ref, gp)                                        # d sum_b(gan + lambda*const_b)
    assert abs(float(gan) - float(gan_ref)) <= 1e-4
    from gan_b200.utils import ssim
    assert np.allclose(np.asarray(total), float(gan_ref) + 100.0 * ssim(x, y), rtol=1e-5, atol=1e-5)
    # SSIM of identical images is 1 (sanity of the host restatement of tf.image.ssim)
    assert np.allclose(ssim(x, x), 1.0)
    worst = 0.0
    for v, r in zip(m.generator.trainable_variables, ref_g):
        r = r.numpy(); den = np.abs(r).max()
        if den > 0:
            worst = max(worst, float(np.abs(v.grad() - r).max() / den))
    assert worst < 5e-2, worst              # batch-2 fp32 conditioning level of the adversarial-only gradient
    m.ctx.close()
