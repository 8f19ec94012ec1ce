"""Host-side mirror of the reference's ``CycleGAN`` (reference cycle_gan.py:26-376), hot path only."""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import _ffi, input_pipeline
from .base_gan import GAN, LossValue, _as_f32
from .utils import cyclegan_losses, save_panel


class CycleGAN(GAN):
    def __init__(self, config):
        """Reference cycle_gan.py:28-37: two InstanceNorm generators, two target=False PatchGANs, four Adams."""
        super().__init__(config)
        ch = int(self.config['channels'])
        self.generator_g = super().Generator(norm_type='instancenorm', shape=(None, None, ch))
        self.generator_f = super().Generator(norm_type='instancenorm', shape=(None, None, ch))
        self.discriminator_x = super().Discriminator(norm_type='instancenorm', target=False)
        self.discriminator_y = super().Discriminator(norm_type='instancenorm', target=False)
        kw = dict(learning_rate=self.config.get('learning_rate', 2e-4), beta_1=self.config.get('beta_1', 0.5),
                  beta_2=self.config.get('beta_2', 0.999))
        self.generator_g_optimizer = super().optimizer(**kw)
        self.generator_f_optimizer = super().optimizer(**kw)
        self.discriminator_x_optimizer = super().optimizer(**kw)
        self.discriminator_y_optimizer = super().optimizer(**kw)

    def generator_loss(self, generated):
        """Reference cycle_gan.py:154-159 (host arrays)."""
        return self.loss_obj(np.ones_like(generated), generated)

    def calc_cycle_loss(self, real_image, cycled_image):
        """Reference cycle_gan.py:161-168."""
        return float(np.mean(np.abs(np.asarray(real_image, np.float64) - np.asarray(cycled_image, np.float64)))) \
            * self.config.get('lambda', 10)

    def identity_loss(self, real_image, same_image):
        """Reference cycle_gan.py:170-177."""
        return self.config.get('lambda', 10) * 0.5 * float(
            np.mean(np.abs(np.asarray(real_image, np.float64) - np.asarray(same_image, np.float64))))

    def train_step(self, real_x, real_y, training: bool = True, sync: bool = True):
        """Reference cycle_gan.py:206-276: six generator and four discriminator forwards, seven
        losses, one fused backward sweep that yields the reference's four gradient sets, four
        Keras-Adam updates.  Returns the seven losses in the reference's order."""
        x, y = _as_f32(real_x), _as_f32(real_y)
        b = int(x.shape[0])
        h = [self.generator_g_optimizer.bind(self.generator_g), self.generator_f_optimizer.bind(self.generator_f),
             self.discriminator_x_optimizer.bind(self.discriminator_x),
             self.discriminator_y_optimizer.bind(self.discriminator_y)]
        losses = np.zeros(7, dtype=np.float32) if sync else None
        _ffi.check(_ffi.lib().gan_cyclegan_train_step(
            self.generator_g.handle, self.generator_f.handle, self.discriminator_x.handle, self.discriminator_y.handle,
            h[0], h[1], h[2], h[3], _ffi.ptr_of(x), _ffi.ptr_of(y), b, C.c_float(float(self.config.get('lambda', 10))),
            int(bool(training)), _ffi.ptr_of(losses)))
        if not sync:
            return None
        return tuple(LossValue(v) for v in losses)

    # -- input pipeline on the device (reference cycle_gan.py:38-85) --------------------------------
    def process_images(self, images, train: bool, rng=None, out=None):
        """``process_images_train`` / ``process_images_pred`` (cycle_gan.py:64-85) for a batch of decoded
        uint8 images: ``load(resize=True)`` resizes to img_size first (base_gan.py:41-43), then
        img_size+30 + random crop + random mirror when ``train`` (independent draws per image,
        cycle_gan.py:46-62), normalize.  One gather kernel on the device."""
        s, c = int(self.config['img_size']), int(self.config['channels'])
        return input_pipeline.preprocess(self.ctx, input_pipeline.pack_images(list(images)), self._xforms(images, train, rng), c, s, out)

    def _xforms(self, images, train, rng):
        s = int(self.config['img_size'])
        xfs = []
        for im in images:
            crop, flip = (0, 0), False
            if train:
                cy, cx, flip = input_pipeline.draw_jitter(rng, s)
                crop = (cy, cx)
            xfs.append(input_pipeline.xform(im.shape[0], im.shape[1], 0, im.shape[1], s, train, crop, flip, pre=s))
        return xfs

    def prefetch_images(self, packed_x, images_x, packed_y, images_y, train: bool, rng=None):
        """Prefetching form: device batches (real_x, real_y) for the next ``train_step``."""
        s, c = int(self.config['img_size']), int(self.config['channels'])
        return input_pipeline.prefetch(self.ctx, packed_x[0], packed_x[1], self._xforms(images_x, train, rng),
                                       packed_y[0], packed_y[1], self._xforms(images_y, train, rng), c, s)

    def image_pipeline(self, predict: bool = False):
        """Reference cycle_gan.py:87-152: list the two unpaired directories, the seeded test/val/train split
        (``random.seed(seed)``; test from X only; val sizes ceil(n*validation_size)), per-iteration shuffling of the
        train and val sets (tf.data ``shuffle(buffer_size, reshuffle_each_iteration=True)``: the order itself is not
        reproducible across frameworks), batching without drop_remainder.  Decoding is PIL on the host, every pixel
        operation runs on the device.  Returns (train_X, train_Y, val_X, val_Y, test)."""
        import random
        cfg = self.config
        contents_X = input_pipeline.list_images(cfg['input_images'])
        assert contents_X, "No images found in input image directory!"
        full = lambda d, names: [os.path.join(d, i) for i in names]   # noqa: E731
        if predict:
            return self._batches(full(cfg['input_images'], contents_X), 1, False, False, squeeze=True), None, None, None, None
        contents_Y = input_pipeline.list_images(cfg['target_images'])
        assert contents_Y, "No images found in target image directory!"
        random.seed(cfg['seed'])
        test = random.sample(contents_X, cfg['test_img'])
        val_obs_X = np.ceil((len(contents_X) - cfg['test_img']) * cfg['validation_size'])
        val_obs_Y = np.ceil(len(contents_Y) * cfg['validation_size'])
        val_X = random.sample([i for i in contents_X if i not in test], int(val_obs_X))
        val_Y = random.sample([i for i in contents_Y], int(val_obs_Y))
        train_X = [i for i in contents_X if i not in test and i not in val_X]
        train_Y = [i for i in contents_Y if i not in val_Y]
        bs, dx, dy = cfg['batch_size'], cfg['input_images'], cfg['target_images']
        return (self._batches(full(dx, train_X), bs, True, True), self._batches(full(dy, train_Y), bs, True, True),
                self._batches(full(dx, val_X), bs, False, True), self._batches(full(dy, val_Y), bs, False, True),
                self._batches(full(dx, test), bs, False, False))

    def _batches(self, files, batch_size, train, shuffle, squeeze=False):
        rng = np.random.default_rng(int(self.config.get('seed', 123)) + 2)
        c = int(self.config['channels'])

        class _DS:
            def __iter__(ds):
                order = list(rng.permutation(len(files))) if shuffle else list(range(len(files)))
                for i in range(0, len(order), batch_size):
                    ims = [input_pipeline.load(files[j], c) for j in order[i:i + batch_size]]
                    a = self.process_images(ims, train, rng)
                    yield a[0] if squeeze else a

            def __len__(ds):
                return (len(files) + batch_size - 1) // batch_size
        return _DS()

    def generate_images(self, model, test_input, path_filename: str = None):
        """Reference cycle_gan.py:179-204: ``model(test_input, training=True)`` and the two-panel figure (input,
        prediction) written as a PNG panel through PIL (or .npy for another suffix)."""
        prediction = model(test_input, training=True)
        if path_filename:
            if path_filename.lower().endswith(".png"):
                save_panel(path_filename, [np.asarray(test_input)[0], prediction[0]], int(self.config['channels']))
            else:
                np.save(path_filename, prediction)
        return prediction

    def fit(self, train_X, train_Y, val_X, val_Y, test=None, output_path: str = None, checkpoint_manager=None):
        """Reference cycle_gan.py:278-358 (zip of the X and Y batch iterables)."""
        example = next(iter(test), None) if (test is not None and output_path) else None     # cycle_gan.py:283
        start = time.time()
        train_cost_functions, val_cost_functions = cyclegan_losses(), cyclegan_losses()
        keys = list(train_cost_functions.keys())
        for epoch in range(self.config['epochs']):
            train_losses, val_losses = cyclegan_losses(), cyclegan_losses()
            for image_x, image_y in zip(train_X, train_Y):
                for k, v in zip(keys, self.train_step(image_x, image_y)):
                    train_losses[k].append(v.numpy().tolist())
            for k in keys:
                train_cost_functions[k].append(sum(train_losses[k]) / len(train_losses[k]))
            for image_x, image_y in zip(val_X, val_Y):
                for k, v in zip(keys, self.train_step(image_x, image_y, False)):
                    val_losses[k].append(v.numpy().tolist())
            for k in keys:
                val_cost_functions[k].append(sum(val_losses[k]) / len(val_losses[k]))
            # every 5 epochs and at the last one: save weights (reference cycle_gan.py:342-350)
            last = (epoch + 1) == self.config['epochs']
            if checkpoint_manager is not None and (((epoch + 1) % 5 == 0) or last):
                checkpoint_manager.save()
            if example is not None and (epoch + 1) % 5 == 0 and not last and self.ctx.rank == 0:
                self.generate_images(self.generator_g, np.asarray(example)[:1],
                                     path_filename=os.path.join(output_path, 'test_images', f"epoch_{epoch + 1}.png"))
            print(f'\nCumulative training duration at end of epoch {epoch + 1}: {(time.time() - start) / 60:.2f} min')
        return train_cost_functions, val_cost_functions

    def predict(self, predict_ds, output_path: str = None):
        """Reference cycle_gan.py:360-376."""
        return [self.generate_images(self.generator_g, np.expand_dims(np.asarray(i), axis=0)) for i in predict_ds]
