"""Host-side mirror of the reference's ``Pix2Pix`` (reference pix2pix.py:26-339), hot path only:
``__init__``, ``generator_loss``, ``train_step``, the forward call of ``generate_images`` /
``predict`` and the ``fit`` epoch loop.  The input pipeline, plotting and TF checkpoints are out
of scope."""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import _ffi
from .base_gan import GAN, LossValue, _as_f32
from .utils import pix2pix_losses


class Pix2Pix(GAN):
    def __init__(self, config):
        """Reference pix2pix.py:27-32: BatchNorm U-Net generator, target=True PatchGAN, two Adams."""
        super().__init__(config)
        size, ch = self.config['img_size'], int(self.config['channels'])
        self.generator = super().Generator(shape=(size, size, ch))
        self.discriminator = super().Discriminator(target=True)
        kw = dict(learning_rate=self.config.get('learning_rate', 2e-4), beta_1=self.config.get('beta_1', 0.5),
                  beta_2=self.config.get('beta_2', 0.999))
        self.generator_optimizer = super().optimizer(**kw)
        self.discriminator_optimizer = super().optimizer(**kw)

    def generator_loss(self, disc_generated_output, gen_output, target, input_image):
        """Reference pix2pix.py:167-188 on host arrays (default 'l1' branch).  ``train_step`` fuses
        the same arithmetic on the device; this method exists for API parity."""
        gan_loss = self.loss_obj(np.ones_like(disc_generated_output), disc_generated_output)
        if self.config.get('generator_loss', 'l1') != 'l1':
            raise NotImplementedError("only the default generator_loss='l1' is on the accelerated path")
        gan_loss2 = float(np.mean(np.abs(np.asarray(target, np.float64) - np.asarray(gen_output, np.float64))))
        return gan_loss + self.config.get('lambda', 100) * gan_loss2, gan_loss, gan_loss2

    def train_step(self, input_image, target, training: bool = True, sync: bool = True):
        """Reference pix2pix.py:190-218.  One fused device step: G forward, D(real), D(fake), the four
        losses, both backward sweeps from the shared forward, both Keras-Adam updates.

        input_image / target: (B,H,W,C) float32 in [-1,1]; numpy (host) or torch tensors (host or
        CUDA).  Returns (gen_total_loss, gen_gan_loss, gen_gan_loss2, disc_loss) as LossValue
        (``.numpy()`` works like on tf scalars).  ``sync=False`` only enqueues the step and returns
        None; read the losses later with ``self.ctx.last_losses(4)``."""
        if self.config.get('generator_loss', 'l1') != 'l1':
            raise NotImplementedError("only the default generator_loss='l1' is on the accelerated path")
        x, y = _as_f32(input_image), _as_f32(target)
        if tuple(x.shape) != tuple(y.shape):
            raise ValueError("input_image and target must have the same shape")
        b = int(x.shape[0])
        g_opt = self.generator_optimizer.bind(self.generator)
        d_opt = self.discriminator_optimizer.bind(self.discriminator)
        losses = np.zeros(4, dtype=np.float32) if sync else None
        _ffi.check(_ffi.lib().gan_pix2pix_train_step(
            self.generator.handle, self.discriminator.handle, g_opt, d_opt, _ffi.ptr_of(x), _ffi.ptr_of(y), b,
            C.c_float(float(self.config.get('lambda', 100))), int(bool(training)), _ffi.ptr_of(losses)))
        if not sync:
            return None
        return tuple(LossValue(v) for v in losses)

    def generate_images(self, model, test_input, tar=None, path_filename: str = None):
        """Forward call of reference pix2pix.py:220-228 (``model(test_input, training=True)``);
        the matplotlib rendering is out of scope — the prediction is returned (and saved as .npy
        when a filename is given)."""
        prediction = model(test_input, training=True)
        if path_filename:
            np.save(path_filename, prediction)
        return prediction

    def fit(self, train_ds, val_ds, test_ds=None, output_path: str = None, checkpoint_manager=None):
        """Reference pix2pix.py:248-323: per-epoch mean of each loss over mini-batches, validation
        through ``train_step(..., False)``.  Datasets are iterables of (input, target) batches."""
        start = time.time()
        train_cost_functions, val_cost_functions = pix2pix_losses(), pix2pix_losses()
        keys = list(train_cost_functions.keys())
        for epoch in range(self.config['epochs']):
            train_losses, val_losses = pix2pix_losses(), pix2pix_losses()
            for input_image, target in train_ds:
                for k, v in zip(keys, self.train_step(input_image, target, True)):
                    train_losses[k].append(v.numpy().tolist())
            for k in keys:
                train_cost_functions[k].append(sum(train_losses[k]) / len(train_losses[k]))
            for input_image, target in val_ds:
                for k, v in zip(keys, self.train_step(input_image, target, False)):
                    val_losses[k].append(v.numpy().tolist())
            for k in keys:
                val_cost_functions[k].append(sum(val_losses[k]) / len(val_losses[k]))
            print(f'\nCumulative training duration at end of epoch {epoch + 1}: {(time.time() - start) / 60:.2f} min')
        return train_cost_functions, val_cost_functions

    def predict(self, predict_ds, output_path: str = None):
        """Reference pix2pix.py:325-339: batch-1 generator forward per (input, target) pair."""
        outs = []
        for i in predict_ds:
            outs.append(self.generate_images(self.generator, np.expand_dims(np.asarray(i[0]), axis=0)))
        return outs
