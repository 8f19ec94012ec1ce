"""Host-side mirror of the reference's ``Pix2Pix`` (reference pix2pix.py:26-339), hot path only:
``__init__``, ``generator_loss``, ``train_step``, the forward call of ``generate_images`` /
``predict``, the ``fit`` epoch loop with its checkpoint cadence, and the input pipeline with every
pixel operation on the device.  Plotting and the TF tensor-bundle format are out of scope (SURVEY.md §2
rows 8, 11, 12); ``gan_b200.checkpoint`` is the own weight/optimizer format."""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import _ffi, input_pipeline
from .base_gan import GAN, LossValue, _as_f32
from .utils import pix2pix_losses, save_panel, ssim


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
        if self.config.get('generator_loss', 'l1') == 'l1':
            gan_loss2 = float(np.mean(np.abs(np.asarray(target, np.float64) - np.asarray(gen_output, np.float64))))
        else:
            # 'ssim' exactly as the reference wrote it (pix2pix.py:182-184): SSIM of the INPUT against the target with
            # max_val=255 on [-1,1] data — a per-image VECTOR that does not depend on the generator
            gan_loss2 = ssim(input_image, target, max_val=255, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03)
        return gan_loss + self.config.get('lambda', 100) * gan_loss2, gan_loss, gan_loss2

    def train_step(self, input_image, target, training: bool = True, sync: bool = True):
        """Reference pix2pix.py:190-218.  One fused device step: G forward, D(real), D(fake), the four
        losses, both backward sweeps from the shared forward, both Keras-Adam updates.

        input_image / target: (B,H,W,C) float32 in [-1,1]; numpy (host) or torch tensors (host or
        CUDA).  Returns (gen_total_loss, gen_gan_loss, gen_gan_loss2, disc_loss) as LossValue
        (``.numpy()`` works like on tf scalars).  ``sync=False`` only enqueues the step and returns
        None; read the losses later with ``self.ctx.last_losses(4)``."""
        x, y = _as_f32(input_image), _as_f32(target)
        if tuple(x.shape) != tuple(y.shape):
            raise ValueError("input_image and target must have the same shape")
        b = int(x.shape[0])
        g_opt = self.generator_optimizer.bind(self.generator)
        d_opt = self.discriminator_optimizer.bind(self.discriminator)
        losses = np.zeros(4, dtype=np.float32) if sync else None
        lam = float(self.config.get('lambda', 100))
        use_ssim = self.config.get('generator_loss', 'l1') != 'l1'
        # 'ssim' (pix2pix.py:182-186): gan_loss2 = SSIM(input, target) is a constant per-image vector, so the generator
        # gradient is that of sum_b(gan_loss + lambda*ssim_b) = batch * d(gan_loss): no L1 term, adversarial term x batch
        l1_weight, gan_scale = (0.0, float(b)) if use_ssim else (lam, 1.0)
        _ffi.check(_ffi.lib().gan_pix2pix_train_step_ex(
            self.generator.handle, self.discriminator.handle, g_opt, d_opt, _ffi.ptr_of(x), _ffi.ptr_of(y), b,
            C.c_float(l1_weight), C.c_float(gan_scale), int(bool(training)), _ffi.ptr_of(losses)))
        if not sync:
            return None
        if use_ssim:
            if hasattr(x, "cpu"):
                x, y = x.cpu().numpy(), y.cpu().numpy()
            s = ssim(x, y, max_val=255, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03).astype(np.float32)
            # the reference returns VECTORS for the total and the secondary loss in this branch
            return (float(losses[1]) + np.float32(lam) * s, LossValue(losses[1]), s, LossValue(losses[3]))
        return tuple(LossValue(v) for v in losses)

    # -- input pipeline on the device (reference pix2pix.py:34-112) ---------------------------------
    def split_img(self, image):
        """Reference pix2pix.py:34-55 on a decoded (H, 2W, C) uint8 pair image (or a file name):
        returns the two column windows as ((col0, cols), (col0, cols)) = (input, real)."""
        if isinstance(image, str):
            image = input_pipeline.load(image, int(self.config['channels']))
        w = image.shape[1] // 2
        left, right = (0, w), (w, image.shape[1] - w)
        return (left, right) if self.config.get('input_img_orient', 'left') == 'left' else (right, left)

    def process_images(self, pairs, train: bool, rng=None, out=None):
        """``process_images_train`` / ``process_images_pred`` (pix2pix.py:92-112) for a batch of decoded
        uint8 pair images: split, nearest resize (to img_size+30 and random crop + mirror when
        ``train``), normalize — one gather kernel per output on the device.  Returns
        (input_image, real_image) as (B, S, S, C) float32 arrays."""
        s, c = int(self.config['img_size']), int(self.config['channels'])
        xa, xb = self._pair_xforms(pairs, train, rng)
        packed = input_pipeline.pack_images(pairs)
        a = input_pipeline.preprocess(self.ctx, packed, xa, c, s, None if out is None else out[0])
        b = input_pipeline.preprocess(self.ctx, packed, xb, c, s, None if out is None else out[1])
        return a, b

    def _pair_xforms(self, pairs, train, rng):
        s = int(self.config['img_size'])
        xa, xb = [], []
        for im in pairs:
            (c0a, na), (c0b, nb) = self.split_img(im)
            crop, flip = (0, 0), False
            if train:
                cy, cx, flip = input_pipeline.draw_jitter(rng, s)     # one draw per PAIR (pix2pix.py:57-69,80-88)
                crop = (cy, cx)
            xa.append(input_pipeline.xform(im.shape[0], im.shape[1], c0a, na, s, train, crop, flip))
            xb.append(input_pipeline.xform(im.shape[0], im.shape[1], c0b, nb, s, train, crop, flip))
        return xa, xb

    def prefetch_pairs(self, packed, pairs_meta, train: bool, rng=None):
        """Prefetching form for the training loop: ``packed`` = (uint8 buffer, stride) from
        ``input_pipeline.pack_images`` (pin it), ``pairs_meta`` = the decoded images (only shapes are
        used).  Returns two device batches for the next ``train_step``."""
        s, c = int(self.config['img_size']), int(self.config['channels'])
        xa, xb = self._pair_xforms(pairs_meta, train, rng)
        buf, stride = packed
        return input_pipeline.prefetch(self.ctx, buf, stride, xa, buf, stride, xb, c, s)

    def image_pipeline(self, predict: bool = False):
        """Reference pix2pix.py:114-165: file listing, the seeded train/val/test split and batching;
        decoding is PIL on the host, every pixel operation runs on the device.  Yields (input,
        target) float32 batches."""
        import random
        cfg = self.config
        contents = input_pipeline.list_images(cfg['data'])
        assert contents, "No images found in data directory!"
        full = lambda names: [os.path.join(cfg['data'], i) for i in names]   # noqa: E731
        if predict:
            return self._batches(full(contents), 1, False, squeeze=True), None, None
        random.seed(cfg['seed'])
        test = random.sample(contents, cfg['test_img'])
        val_obs = np.ceil((len(contents) - cfg['test_img']) * cfg['validation_size'])
        val = random.sample([i for i in contents if i not in test], int(val_obs))
        train = [i for i in contents if i not in test and i not in val]
        train = random.sample(train, len(train))
        bs = cfg['batch_size']
        return (self._batches(full(train), bs, True), self._batches(full(val), bs, False),
                self._batches(full(test), bs, False))

    def _batches(self, files, batch_size, train, squeeze=False):
        rng = np.random.default_rng(int(self.config.get('seed', 123)) + 2)
        c = int(self.config['channels'])

        class _DS:
            def __iter__(ds):
                for i in range(0, len(files), batch_size):
                    pairs = [input_pipeline.load(f, c) for f in files[i:i + batch_size]]
                    a, b = self.process_images(pairs, train, rng)
                    yield (a[0], b[0]) if squeeze else (a, b)

            def __len__(ds):
                return (len(files) + batch_size - 1) // batch_size
        return _DS()

    def generate_images(self, model, test_input, tar=None, path_filename: str = None):
        """Reference pix2pix.py:220-246: ``model(test_input, training=True)`` and the three-panel figure
        (input, ground truth, prediction; values x*0.5+0.5) written to ``path_filename`` — as a PNG panel through PIL
        (matplotlib is not a dependency), or the raw prediction as .npy for any other suffix.  The prediction is
        also returned."""
        prediction = model(test_input, training=True)
        if path_filename:
            if path_filename.lower().endswith(".png"):
                panels = [np.asarray(test_input)[0]] + ([np.asarray(tar)[0]] if tar is not None else []) + [prediction[0]]
                save_panel(path_filename, panels, int(self.config['channels']))
            else:
                np.save(path_filename, prediction)
        return prediction

    def fit(self, train_ds, val_ds, test_ds=None, output_path: str = None, checkpoint_manager=None):
        """Reference pix2pix.py:248-323: per-epoch mean of each loss over mini-batches, validation
        through ``train_step(..., False)``.  Datasets are iterables of (input, target) batches."""
        example = None
        if test_ds is not None and output_path:
            example = next(iter(test_ds), None)            # example_input, example_target (pix2pix.py:260)
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
            # every 5 epochs and at the last one: save weights (reference pix2pix.py:308-317)
            last = (epoch + 1) == self.config['epochs']
            if checkpoint_manager is not None and (((epoch + 1) % 5 == 0) or last):
                checkpoint_manager.save()
            # every 5 epochs (not the last): predicted image of the first test example (pix2pix.py:307-313)
            if example is not None and (epoch + 1) % 5 == 0 and not last and self.ctx.rank == 0:
                self.generate_images(self.generator, np.asarray(example[0])[:1], np.asarray(example[1])[:1],
                                     path_filename=os.path.join(output_path, 'test_images', f"epoch_{epoch + 1}.png"))
            print(f'\nCumulative training duration at end of epoch {epoch + 1}: {(time.time() - start) / 60:.2f} min')
        return train_cost_functions, val_cost_functions

    def predict(self, predict_ds, output_path: str = None):
        """Reference pix2pix.py:325-339: batch-1 generator forward per (input, target) pair; with ``output_path``
        the panels are written to ``<output_path>/prediction_images/img<k>.png`` as the reference does."""
        outs = []
        for k, i in enumerate(predict_ds):
            fn = os.path.join(output_path, 'prediction_images', f"img{k}.png") if output_path else None
            outs.append(self.generate_images(self.generator, np.expand_dims(np.asarray(i[0]), axis=0),
                                             np.expand_dims(np.asarray(i[1]), axis=0), fn))
        return outs
