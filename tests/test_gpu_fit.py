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
               precision='bf16', epochs=5, batch_size=2, data=str(data), test_img=1, validation_size=0.2,
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
    assert list(tr.keys()) == keys and all(len(tr[k]) == 5 and len(va[k]) == 5 for k in keys)
    assert all(np.isfinite(vv) for k in keys for vv in tr[k] + va[k])
    assert m.generator_optimizer.iterations == 5 * 2
    assert os.path.basename(latest_checkpoint(str(ckdir))) == "ckpt-1.npz"     # epoch 5 == last epoch: one save

    pred_ds, _, _ = m.image_pipeline(predict=True)
    outs = m.predict(pred_ds)
    assert len(outs) == 6 and outs[0].shape == (1, 256, 256, 3) and np.abs(outs[0]).max() <= 1.0
    m.ctx.close()
