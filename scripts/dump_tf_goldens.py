#!/usr/bin/env python
"""Pin the oracle against the REAL reference: dump golden vectors from kingjosephm/GAN under TensorFlow 2.6.

This container cannot run it (TensorFlow is not installable here: no network, no cp312 wheel; see DESIGN.md
§5), so it is the hand-off half of the bridge: a TF-2.6 user runs

    PYTHONPATH=/path/to/kingjosephm-GAN:/path/to/this/repo \
        python scripts/dump_tf_goldens.py --out tests/golden/tf --model pix2pix --batch 2 --steps 3

and commits the resulting ``tests/golden/tf/*.npz``; ``tests/test_tf_goldens.py`` (skipped while the directory is
empty) then checks the CPU oracle — and through it every device parity test — against TensorFlow's own arithmetic:
losses, every gradient tensor, weights and Adam slots after each step.

What it does, and what it deliberately does NOT touch:
  * imports the reference's own ``pix2pix.Pix2Pix`` / ``cycle_gan.CycleGAN`` classes (base_gan.py:124-252,
    pix2pix.py:190-218, cycle_gan.py:206-276) and calls their unmodified ``train_step`` (the @tf.function);
  * weights: generated on the host with ``oracle.gan_oracle.init_params(default_rng(seed+1))`` and assigned in Keras
    ``trainable_variables`` order — the variable NAMES and SHAPES TensorFlow reports are stored
    too, so a mismatch of that order is visible in the file;
  * inputs: ``oracle.gan_oracle.synthetic_images(default_rng(seed))``;
  * dropout: TensorFlow's mask stream cannot be reproduced elsewhere, so ``tf.keras.layers.Dropout`` is replaced
    BEFORE the reference modules are imported by a layer that multiplies with an explicit keep mask held in a
    ``tf.Variable`` (kept values scaled by 1/(1-rate), exactly Dropout's training behaviour); the masks are the
    oracle's Philox masks keyed (seed, generator-call counter, layer tag 1..3, sample, element) and are stored in
    the file.  Nothing else of TensorFlow or of the reference is altered.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def install_mask_dropout(tf):
    """Replace tf.keras.layers.Dropout by an explicit-mask layer; returns the registry of created layers."""
    created = []

    class MaskDropout(tf.keras.layers.Layer):
        def __init__(self, rate, **kw):
            super().__init__(**kw)
            self.rate = float(rate)
            self.mask = None
            created.append(self)

        def build(self, input_shape):
            self._feat = tuple(int(d) for d in input_shape[1:])

        def set_mask(self, keep):                      # keep: (B, H, W, C) {0,1} float32
            if self.mask is None or tuple(self.mask.shape) != tuple(keep.shape):
                self.mask = tf.Variable(keep, trainable=False, dtype=tf.float32)
            else:
                self.mask.assign(keep)

        def call(self, x, training=None):
            if self.mask is None:
                raise RuntimeError("MaskDropout: set_mask() before the first call")
            return x * self.mask * (1.0 / (1.0 - self.rate))

    tf.keras.layers.Dropout = MaskDropout
    return created


def set_weights(model, arrays):
    tv = model.trainable_variables
    assert len(tv) == len(arrays), (len(tv), len(arrays))
    for v, a in zip(tv, arrays):
        assert tuple(v.shape) == tuple(a.shape), (v.name, v.shape, a.shape)
        v.assign(a)


def adam_slots(opt, model):
    out = {}
    for v in model.trainable_variables:
        out[v.name + "/m"] = opt.get_slot(v, "m").numpy()
        out[v.name + "/v"] = opt.get_slot(v, "v").numpy()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "tf"))
    ap.add_argument("--model", default="pix2pix", choices=["pix2pix", "cyclegan"])
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--channels", type=int, default=3)
    ap.add_argument("--seed", type=int, default=123)
    args = ap.parse_args()

    import tensorflow as tf
    drops = install_mask_dropout(tf)                   # before the reference builds its layers
    from oracle import gan_oracle as O
    B, S, C, seed = args.batch, args.size, args.channels, args.seed
    rng_w = np.random.default_rng(seed + 1)
    irng = np.random.default_rng(seed)
    x = O.synthetic_images(irng, B, S, S, C); y = O.synthetic_images(irng, B, S, S, C)
    out = {"protocol": np.array([seed, B, S, C, args.steps]), "x": x, "y": y,
           "tf_version": np.frombuffer(tf.__version__.encode(), dtype=np.uint8)}

    if args.model == "pix2pix":
        from pix2pix import Pix2Pix                    # the reference's own module
        cfg = {"img_size": S, "channels": str(C), "learning_rate": 2e-4, "beta_1": 0.5, "beta_2": 0.999, "lambda": 100,
               "generator_loss": "l1", "batch_size": B, "seed": seed, "input_img_orient": "left", "epochs": 1}
        m = Pix2Pix(cfg)
        g_np = O.init_params(O.generator_spec(C), rng_w, "batchnorm")
        d_np = O.init_params(O.discriminator_spec(C, True), rng_w, "batchnorm")
        set_weights(m.generator, g_np); set_weights(m.discriminator, d_np)
        nets = {"g": (m.generator, m.generator_optimizer), "d": (m.discriminator, m.discriminator_optimizer)}
        gen_drops = {"g": drops[:3]}
        calls_per_step = ["g"]
    else:
        from cycle_gan import CycleGAN
        cfg = {"img_size": S, "channels": str(C), "learning_rate": 2e-4, "beta_1": 0.5, "beta_2": 0.999, "lambda": 10,
               "batch_size": B, "seed": seed, "epochs": 1}
        m = CycleGAN(cfg)
        specs = [O.generator_spec(C), O.generator_spec(C), O.discriminator_spec(C, False), O.discriminator_spec(C, False)]
        arrs = [O.init_params(s, rng_w, "instancenorm") for s in specs]
        mods = [m.generator_g, m.generator_f, m.discriminator_x, m.discriminator_y]
        opts = [m.generator_g_optimizer, m.generator_f_optimizer, m.discriminator_x_optimizer, m.discriminator_y_optimizer]
        for mod, a in zip(mods, arrs):
            set_weights(mod, a)
        nets = dict(zip(["g", "f", "dx", "dy"], zip(mods, opts)))
        gen_drops = {"g": drops[:3], "f": drops[3:6]}
        # A generator is called three times per step with DIFFERENT batches (cycle_gan.py:220-228) but a Keras layer
        # holds ONE mask variable: dump CycleGAN with dropout disabled (all-ones keep mask, scale compensated by
        # rate=0) unless you extend MaskDropout with a per-call counter.  Recorded in the file as dropout=0.
        for d in drops:
            d.rate = 0.0
        calls_per_step = []

    for tag, (mod, _) in nets.items():
        out[f"{tag}/names"] = np.array([v.name for v in mod.trainable_variables])
        for i, v in enumerate(mod.trainable_variables):
            out[f"{tag}/w0/{i}"] = v.numpy()

    xt, yt = tf.constant(x), tf.constant(y)
    call = 0
    for s in range(args.steps):
        if args.model == "pix2pix":
            shapes = [(B, 2 * S // 256, 2 * S // 256, 512), (B, 4 * S // 256, 4 * S // 256, 512), (B, 8 * S // 256, 8 * S // 256, 512)]
            for tagl, (layer, shp) in enumerate(zip(gen_drops["g"], shapes), start=1):
                keep = O.dropout_keep_mask(seed, call, tagl, 0, shp).astype(np.float32)
                layer.set_mask(keep)
                out[f"step{s}/mask{tagl}"] = keep.astype(np.uint8)
            call += len(calls_per_step)
        else:
            for layer in drops:
                shp = (B,) + tuple(layer._feat) if hasattr(layer, "_feat") else None
                if shp is None:                         # not built yet: build by a dry forward with ones
                    continue
                layer.set_mask(np.ones(shp, np.float32))
            out["dropout"] = np.array(0)
        # gradients of THIS step at the pre-update weights: recompute them the way train_step does, with the tapes
        # (train_step itself returns only the losses); then run the reference's own train_step for the update
        if args.model == "pix2pix":
            with tf.GradientTape() as gt, tf.GradientTape() as dt:
                gen_output = m.generator(xt, training=True)
                disc_real = m.discriminator([xt, yt], training=True)
                disc_fake = m.discriminator([xt, gen_output], training=True)
                total, gan, l1 = m.generator_loss(disc_fake, gen_output, yt, xt)
                dl = m.discriminator_loss(disc_real, disc_fake, 0.5)
            gg = gt.gradient(total, m.generator.trainable_variables)
            dg = dt.gradient(dl, m.discriminator.trainable_variables)
            for i, g in enumerate(gg):
                out[f"step{s}/g/grad/{i}"] = g.numpy()
            for i, g in enumerate(dg):
                out[f"step{s}/d/grad/{i}"] = g.numpy()
            out[f"step{s}/gen_output"] = gen_output.numpy()
            out[f"step{s}/tape_losses"] = np.array([float(total), float(gan), float(l1), float(dl)])
            # the tapes' forward calls moved the BatchNorm moving statistics; they are never read on the path
        losses = m.train_step(xt, yt, True)             # the reference's @tf.function, unmodified
        out[f"step{s}/losses"] = np.array([float(v) for v in losses])
        for tag, (mod, opt) in nets.items():
            for i, v in enumerate(mod.trainable_variables):
                out[f"step{s}/{tag}/w/{i}"] = v.numpy()
            for k, a in adam_slots(opt, mod).items():
                out[f"step{s}/{tag}/slot/{k}"] = a
            out[f"step{s}/{tag}/iterations"] = np.array(int(opt.iterations.numpy()))

    os.makedirs(args.out, exist_ok=True)
    path = os.path.join(args.out, f"{args.model}_b{B}_c{C}_s{S}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
