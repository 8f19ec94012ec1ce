"""Weight / optimizer export-import round trip (gan_b200/checkpoint.py, the role of
tf.train.Checkpoint + CheckpointManager in pix2pix.py:399-422): a restored model continues the
training trajectory bit for bit."""
import os

import numpy as np
import pytest

from oracle import gan_oracle as O

pytestmark = pytest.mark.gpu

SEED = 123


def _build(precision="fp32", seed_shift=0):
    from gan_b200 import Pix2Pix
    cfg = dict(img_size=256, channels='3', learning_rate=2e-4, beta_1=0.5, beta_2=0.999, generator_loss='l1',
               seed=SEED + seed_shift, precision=precision, epochs=1, batch_size=1)
    cfg['lambda'] = 100
    return Pix2Pix(cfg)


def _ckpt(m):
    from gan_b200 import Checkpoint
    return Checkpoint(generator_optimizer=m.generator_optimizer, discriminator_optimizer=m.discriminator_optimizer,
                      generator=m.generator, discriminator=m.discriminator)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_restore_continues_the_trajectory_bitwise(tmp_path, precision):
    from gan_b200 import CheckpointManager
    rng = np.random.default_rng(4)
    x, y = O.synthetic_images(rng, 2, 256, 256, 3), O.synthetic_images(rng, 2, 256, 256, 3)
    a = _build(precision)
    a.ctx.set_graphs(False)          # eager: wgrad reductions are atomics, graph or not; forward is deterministic
    for _ in range(2):
        a.train_step(x, y, True)
    mgr = CheckpointManager(_ckpt(a), str(tmp_path), max_to_keep=1)
    path = mgr.save()
    counter = a.ctx.call_counter()
    assert os.path.basename(path) == "ckpt-1.npz" and mgr.latest_checkpoint == path
    want = [float(v) for v in a.train_step(x, y, False)]              # validation step: pure function of the state

    b = _build(precision, seed_shift=50)                              # different random init: everything must come from the file
    status = _ckpt(b).restore(mgr.latest_checkpoint,
                              models_for_optimizers={"generator_optimizer": b.generator,
                                                     "discriminator_optimizer": b.discriminator})
    status.assert_consumed()
    b.ctx.set_graphs(False)
    b.ctx.set_rng(SEED, counter)
    got = [float(v) for v in b.train_step(x, y, False)]
    assert got == want
    assert b.generator_optimizer.iterations == 2 and b.discriminator_optimizer.iterations == 2
    for va, vb in zip(a.generator.variables + a.discriminator.variables, b.generator.variables + b.discriminator.variables):
        assert va.name == vb.name and np.array_equal(va.numpy(), vb.numpy()), va.name
    for which in ("m", "v"):
        assert np.array_equal(a.generator_optimizer.get_state(which), b.generator_optimizer.get_state(which))
        assert np.array_equal(a.discriminator_optimizer.get_state(which), b.discriminator_optimizer.get_state(which))
    # BatchNorm moving statistics travelled too (two training forwards moved them off their 0 / 1 init)
    mov = [v for v in b.generator.variables if not v.trainable]
    assert mov and any(np.abs(v.numpy()).max() > 0 for v in mov if v.name.endswith("moving_mean"))
    a.ctx.close(); b.ctx.close()


def test_manager_rotation_partial_restore_and_errors(tmp_path):
    from gan_b200 import Checkpoint, CheckpointManager, latest_checkpoint
    m = _build("bf16")
    assert latest_checkpoint(str(tmp_path)) is None
    mgr = CheckpointManager(Checkpoint(generator=m.generator), str(tmp_path), max_to_keep=2)
    paths = [mgr.save() for _ in range(3)]
    assert [os.path.basename(p) for p in mgr.checkpoints] == ["ckpt-2.npz", "ckpt-3.npz"] and not os.path.exists(paths[0])
    assert CheckpointManager(Checkpoint(generator=m.generator), str(tmp_path), 2).save().endswith("ckpt-4.npz")   # resumes numbering
    # predict-style partial restore (pix2pix.py:411 .expect_partial()): the file lacks the discriminator
    st = Checkpoint(generator=m.generator, discriminator=m.discriminator).restore(latest_checkpoint(str(tmp_path)))
    assert st.missing and all(k.startswith("discriminator/") for k in st.missing)
    st.expect_partial()
    with pytest.raises(AssertionError):
        st.assert_consumed()
    # channel mismatch (the reference's comment at pix2pix.py:411): clear error, not silent corruption
    from gan_b200 import Pix2Pix
    cfg = dict(m.config); cfg['channels'] = '1'
    g1 = Pix2Pix(cfg)
    with pytest.raises(ValueError):
        Checkpoint(generator=g1.generator).restore(latest_checkpoint(str(tmp_path)))
    with pytest.raises(TypeError):
        Checkpoint(thing=np.zeros(3))
    m.ctx.close(); g1.ctx.close()


def test_reference_predict_flow_restores_into_unbound_optimizers(tmp_path):
    """The reference's predict path (pix2pix.py:400-411): build a fresh model, wrap models AND optimizers in a
    Checkpoint, restore(latest).expect_partial() BEFORE any train_step — the optimizers have no slots yet.  The
    slots must be deferred (TF's deferred restoration), not an error, and applied when training resumes."""
    from gan_b200 import CheckpointManager
    rng = np.random.default_rng(8)
    x, y = O.synthetic_images(rng, 2, 256, 256, 3), O.synthetic_images(rng, 2, 256, 256, 3)
    a = _build("fp32")
    a.ctx.set_graphs(False)
    for _ in range(2):
        a.train_step(x, y, True)
    mgr = CheckpointManager(_ckpt(a), str(tmp_path), max_to_keep=1)
    mgr.save()
    # a crashed writer's leftover must not break the manager (glob vs regex)
    open(os.path.join(str(tmp_path), "ckpt-9.npz.tmp.npz"), "wb").close()
    assert [os.path.basename(p) for p in mgr.checkpoints] == ["ckpt-1.npz"]
    b = _build("fp32", seed_shift=9)
    b.generator_optimizer.learning_rate = 1e-3                           # differs from the file: the file wins, as in TF
    status = _ckpt(b).restore(mgr.latest_checkpoint)
    status.expect_partial()
    status.assert_consumed()                                             # deferred slots count as consumed
    assert b.generator_optimizer.iterations == 2 and b.generator_optimizer.learning_rate == 2e-4
    out_a, out_b = a.generator(x), None
    b.ctx.set_rng(SEED, a.ctx.call_counter() - 1)
    out_b = b.generator(x)
    assert np.array_equal(out_a, out_b)                                  # predict: weights came from the file
    # training resumes with the restored slots: one more step on both gives identical weights
    b.ctx.set_graphs(False)
    a.ctx.set_rng(SEED, 100); b.ctx.set_rng(SEED, 100)
    a.train_step(x, y, True); b.train_step(x, y, True)
    assert b.generator_optimizer.iterations == 3
    for which in ("m", "v"):
        assert np.allclose(a.generator_optimizer.get_state(which), b.generator_optimizer.get_state(which), rtol=1e-5, atol=1e-7)
    assert np.abs(a.generator.get_flat_params() - b.generator.get_flat_params()).max() <= 2 * 2e-4 * 1.01
    a.ctx.close(); b.ctx.close()
