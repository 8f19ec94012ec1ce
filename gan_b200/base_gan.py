"""Host-side mirror of the reference's ``base_gan.GAN`` (reference base_gan.py:21-292).

Same class / method names and argument meaning as the reference so that ``pix2pix.py`` /
``cycle_gan.py`` style drivers run unchanged; every model call and the whole train step are
executed by hand-written sm_100a kernels behind the C-ABI in ``include/gan_b200.h``.  NumPy is
used for host arrays, torch (optional) only as an allocator of pinned / device buffers and for the
``torch.distributed`` rendezvous that hands the NCCL unique id to the library.
"""
from __future__ import annotations

import ctypes as C
import os
from abc import ABC, abstractmethod

import numpy as np

from . import _ffi

_NORMS = {"batchnorm": _ffi.GAN_NORM_BATCH, "instancenorm": _ffi.GAN_NORM_INSTANCE}
_PRECISIONS = {"fp32": _ffi.GAN_FP32, "bf16": _ffi.GAN_BF16}


def _as_f32(x):
    """Host arrays are normalised to C-contiguous float32; torch tensors pass through."""
    if hasattr(x, "data_ptr"):
        return x
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


class LossValue(float):
    """Scalar returned by train_step; callers do ``.numpy().tolist()`` (reference pix2pix.py:276-279)."""

    def numpy(self):
        return np.float32(self)


class Context:
    """Owns one ``gan_ctx``: the device, stream, workspaces and (optionally) the NCCL communicator."""

    def __init__(self, device: int = 0, precision: str = "bf16", seed: int = 123):
        self.precision = precision
        self.seed = int(seed)
        self._h = C.c_void_p()
        _ffi.check(_ffi.lib().gan_ctx_create(int(device), _PRECISIONS[precision], C.c_uint64(seed), C.byref(self._h)))
        self.device = device
        self.rank, self.world = 0, 1

    @property
    def handle(self):
        return self._h

    def sync(self):
        _ffi.check(_ffi.lib().gan_ctx_sync(self._h))

    def set_dropout(self, enabled: bool):
        _ffi.check(_ffi.lib().gan_ctx_set_dropout(self._h, int(bool(enabled))))

    def set_rng(self, seed: int, call_counter: int = 0):
        self.seed = int(seed)
        _ffi.check(_ffi.lib().gan_ctx_set_rng(self._h, C.c_uint64(seed), C.c_uint32(call_counter)))

    def call_counter(self) -> int:
        v = C.c_uint32()
        _ffi.check(_ffi.lib().gan_ctx_get_call_counter(self._h, C.byref(v)))
        return v.value

    def set_engine(self, engine: int):
        _ffi.check(_ffi.lib().gan_ctx_set_engine(self._h, int(engine)))

    def set_graphs(self, enabled: bool):
        _ffi.check(_ffi.lib().gan_ctx_set_graphs(self._h, int(bool(enabled))))

    def launch_count(self) -> int:
        v = C.c_uint64()
        _ffi.check(_ffi.lib().gan_ctx_launch_count(self._h, C.byref(v)))
        return v.value

    PROFILE_FAMILIES = ("umma_fwd", "umma_wgrad", "ffma_fwd", "ffma_wgrad", "norm", "adam", "pack", "other")

    def set_profile(self, enabled: bool):
        _ffi.check(_ffi.lib().gan_ctx_set_profile(self._h, int(bool(enabled))))

    def profile_read(self):
        """{family: (total ms, algorithmic work, launches)} since the last read (synchronises)."""
        ms, work, cnt = (C.c_double * 8)(), (C.c_double * 8)(), (C.c_int64 * 8)()
        _ffi.check(_ffi.lib().gan_ctx_profile_read(self._h, ms, work, cnt))
        return {n: (ms[i], work[i], cnt[i]) for i, n in enumerate(self.PROFILE_FAMILIES)}

    def stream(self) -> int:
        v = C.c_void_p()
        _ffi.check(_ffi.lib().gan_ctx_stream(self._h, C.byref(v)))
        return v.value or 0

    def last_losses(self, n: int):
        out = np.zeros(n, dtype=np.float32)
        _ffi.check(_ffi.lib().gan_ctx_last_losses(self._h, _ffi.ptr_of(out), n))
        return out

    def prefetch(self, x, y):
        """Start the host->device copy of the next step's (input, target) batches (pinned host arrays /
        tensors) while the current step computes; pass the SAME objects to the next train_step."""
        x, y = _as_f32(x), _as_f32(y)
        nbytes = int(np.prod(x.shape)) * 4
        _ffi.check(_ffi.lib().gan_ctx_prefetch(self._h, _ffi.ptr_of(x), _ffi.ptr_of(y), C.c_int64(nbytes)))
        return x, y

    def set_sample_offset(self, sample0: int):
        _ffi.check(_ffi.lib().gan_ctx_set_sample_offset(self._h, C.c_int64(sample0)))

    def init_data_parallel(self, rank: int, world: int, unique_id: bytes):
        """Join the NCCL communicator; ``unique_id`` is rank 0's ``nccl_unique_id()`` (128 bytes)."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _ffi.check(_ffi.lib().gan_ctx_comm_init(self._h, int(rank), int(world), C.cast(buf, C.c_void_p)))
        self.rank, self.world = rank, world

    def op_conv(self, kind: int, role: int, a, b, batch, height, width, cin, cout, engine=_ffi.GAN_ENGINE_AUTO):
        """Single-operator entry (tests): see gan_op_conv in include/gan_b200.h."""
        a = _as_f32(a); b = _as_f32(b)
        if kind == 0:
            ho, wo = height // 2, width // 2
        elif kind == 1:
            ho, wo = height - 1, width - 1
        else:
            ho, wo = height * 2, width * 2
        if role == 0:
            out = np.empty((batch, ho, wo, cout), dtype=np.float32)
        elif role == 1:
            out = np.empty((batch, height, width, cin), dtype=np.float32)
        else:
            out = np.empty((4, 4, cout, cin) if kind == 2 else (4, 4, cin, cout), dtype=np.float32)
        _ffi.check(_ffi.lib().gan_op_conv(self._h, kind, role, engine, _ffi.ptr_of(a), _ffi.ptr_of(b), _ffi.ptr_of(out),
                                          batch, height, width, cin, cout))
        return out

    def close(self):
        """Destroy the context and every net / optimizer created from it (their handles become invalid)."""
        if self._h:
            _ffi.lib().gan_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001  (interpreter shutdown: the library may already be gone)
            pass


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _ffi.check(_ffi.lib().gan_comm_unique_id(C.cast(buf, C.c_void_p)))
    return buf.raw


def exchange_unique_id(make_id=nccl_unique_id):
    """Rank 0 creates the NCCL unique id, everybody receives it through torch.distributed
    (any backend; plumbing only).  Returns (rank, world, id_bytes)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return rank, world, box[0]


def shard_bounds(global_batch: int, rank: int, world: int):
    """Sample range [lo, hi) of ``rank`` when a global batch is split evenly by sample index
    (SURVEY 8e).  The global batch must divide evenly so that the mean of rank means is the mean."""
    if global_batch % world != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


class Variable:
    """One trainable (or moving-statistic) tensor of a Model, living in device memory."""

    def __init__(self, model, idx, name, shape, trainable):
        self._model, self._idx = model, idx
        self.name, self.shape, self.trainable = name, tuple(shape), trainable

    def numpy(self):
        out = np.empty(self.shape, dtype=np.float32)
        _ffi.check(_ffi.lib().gan_net_get_tensor(self._model._h, self._idx, _ffi.ptr_of(out)))
        return out

    def assign(self, value):
        v = np.ascontiguousarray(np.asarray(value, dtype=np.float32))
        if v.shape != self.shape:
            raise ValueError(f"{self.name}: shape {v.shape} != {self.shape}")
        _ffi.check(_ffi.lib().gan_net_set_tensor(self._model._h, self._idx, _ffi.ptr_of(v)))

    def grad(self):
        out = np.empty(self.shape, dtype=np.float32)
        _ffi.check(_ffi.lib().gan_net_get_grad(self._model._h, self._idx, _ffi.ptr_of(out)))
        return out


class Model:
    """Callable wrapper of one ``gan_net`` — stands in for the ``tf.keras.Model`` returned by
    ``GAN.Generator`` / ``GAN.Discriminator`` (reference base_gan.py:166,225)."""

    def __init__(self, ctx: Context, handle, kind: str, shape, channels: int, target: bool = False):
        self._ctx, self._h, self.kind = ctx, handle, kind
        self.input_hw, self.channels, self.target = shape, channels, target
        nt, tot = C.c_int(), C.c_int()
        _ffi.check(_ffi.lib().gan_net_num_tensors(self._h, C.byref(nt), C.byref(tot)))
        self.variables = []
        name = C.create_string_buffer(64)
        for i in range(tot.value):
            ndim, numel = C.c_int(), C.c_int64()
            shp = (C.c_int64 * 4)()
            _ffi.check(_ffi.lib().gan_net_tensor_info(self._h, i, name, 64, C.byref(ndim), shp, C.byref(numel)))
            self.variables.append(Variable(self, i, name.value.decode(), [shp[k] for k in range(ndim.value)],
                                           i < nt.value))
        self.trainable_variables = [v for v in self.variables if v.trainable]

    @property
    def handle(self):
        return self._h

    def num_params(self) -> int:
        n = C.c_int64()
        _ffi.check(_ffi.lib().gan_net_num_params(self._h, C.byref(n)))
        return n.value

    def get_weights(self):
        return [v.numpy() for v in self.trainable_variables]

    def set_weights(self, arrays):
        for v, a in zip(self.trainable_variables, arrays):
            v.assign(a)

    def get_flat_params(self):
        out = np.empty(self.num_params(), dtype=np.float32)
        _ffi.check(_ffi.lib().gan_net_get_params(self._h, _ffi.ptr_of(out)))
        return out

    def get_flat_grads(self):
        out = np.empty(self.num_params(), dtype=np.float32)
        _ffi.check(_ffi.lib().gan_net_get_grads(self._h, _ffi.ptr_of(out)))
        return out

    def debug_tensor(self, name: str, slot: int = 0, cap: int = 1 << 28):
        n = C.c_int64()
        buf = np.empty(cap // 4, dtype=np.float32)
        _ffi.check(_ffi.lib().gan_net_debug_tensor(self._h, slot, name.encode(), _ffi.ptr_of(buf), buf.size, C.byref(n)))
        return buf[:n.value].copy()

    def __call__(self, inputs, training=True, out=None):
        """``model(x, training=True)`` (reference pix2pix.py:200-203,228).  As in the reference every
        call uses batch statistics and active dropout; ``training`` is accepted and ignored the same
        way Keras ignores it here (the reference never passes False)."""
        if self.kind == "generator":
            x = _as_f32(inputs)
            b, h, w, c = x.shape
            if (h, w) != tuple(self.input_hw) or c != self.channels:
                raise ValueError(f"generator built for {self.input_hw}x{self.channels}, got {(h, w, c)}")
            if out is None:
                out = np.empty((b, h, w, c), dtype=np.float32)
            _ffi.check(_ffi.lib().gan_generator_forward(self._h, _ffi.ptr_of(x), b, _ffi.ptr_of(out)))
            return out
        if self.target:
            inp, tar = inputs
            inp, tar = _as_f32(inp), _as_f32(tar)
        else:
            inp, tar = _as_f32(inputs), None
        b, h, w, _ = inp.shape
        if out is None:
            out = np.empty((b, h // 8 - 2, w // 8 - 2, 1), dtype=np.float32)
        _ffi.check(_ffi.lib().gan_discriminator_forward(self._h, _ffi.ptr_of(inp), _ffi.ptr_of(tar), b, h, w,
                                                        _ffi.ptr_of(out)))
        return out


class Adam:
    """``tf.keras.optimizers.Adam(learning_rate, beta_1, beta_2)`` (reference base_gan.py:247-252):
    epsilon 1e-7 outside the bias correction, one step counter per optimizer.  Bound to its model
    at first use, as Keras creates slots at the first ``apply_gradients``."""

    def __init__(self, learning_rate=2e-4, beta_1=0.5, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self._h = None
        self._model = None
        self._deferred = None      # checkpoint state restored before the slots existed (applied by bind)

    def bind(self, model: Model):
        if self._h is None:
            h = C.c_void_p()
            _ffi.check(_ffi.lib().gan_adam_create(model.handle, self.learning_rate, self.beta_1, self.beta_2,
                                                  self.epsilon, C.byref(h)))
            self._h, self._model = h, model
            if self._deferred is not None:
                state, self._deferred = self._deferred, None
                self._apply_state(state)
        elif self._model is not model:
            raise ValueError("optimizer already bound to another model")
        return self._h

    def set_hyper(self, learning_rate, beta_1, beta_2, epsilon):
        """Replace the hyper-parameters (a restored checkpoint carries its own, as a TF checkpoint does)."""
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        if self._h is not None:
            _ffi.check(_ffi.lib().gan_adam_set_hyper(self._h, learning_rate, beta_1, beta_2, epsilon))

    def _apply_state(self, state):
        """state = {'iterations': t, 'slots': {'m/<var>': array, 'v/<var>': array}}; missing slots stay zero."""
        for i, which in enumerate(("m", "v")):
            parts = []
            for var in self._model.trainable_variables:
                a = state["slots"].get(f"{which}/{var.name}")
                if a is not None and a.shape != var.shape:
                    raise ValueError(f"{which}/{var.name}: shape {a.shape} != {var.shape}")
                parts.append(np.zeros(int(np.prod(var.shape)), np.float32) if a is None
                             else np.asarray(a, np.float32).reshape(-1))
            flat = np.ascontiguousarray(np.concatenate(parts))
            _ffi.check(_ffi.lib().gan_adam_set_state(self._h, i, _ffi.ptr_of(flat)))
        _ffi.check(_ffi.lib().gan_adam_set_step(self._h, C.c_int64(int(state["iterations"]))))

    @property
    def iterations(self) -> int:
        if self._h is None:
            return int(self._deferred["iterations"]) if self._deferred is not None else 0
        t = C.c_int64()
        _ffi.check(_ffi.lib().gan_adam_get_step(self._h, C.byref(t)))
        return t.value

    def get_state(self, which: str):
        out = np.empty(self._model.num_params(), dtype=np.float32)
        _ffi.check(_ffi.lib().gan_adam_get_state(self._h, 0 if which == "m" else 1, _ffi.ptr_of(out)))
        return out


class BinaryCrossentropyFromLogits:
    """Host-side ``tf.keras.losses.BinaryCrossentropy(from_logits=True)`` for callers that evaluate
    a loss on arrays they already hold (reference base_gan.py:227-231).  ``train_step`` does not
    use this object: the adversarial losses are fused into the device step."""

    def __call__(self, y_true, y_pred):
        x = np.asarray(y_pred, dtype=np.float64); z = np.asarray(y_true, dtype=np.float64)
        return float(np.mean(np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))))


class GAN(ABC):
    """Reference base_gan.py:21-24.  ``config`` takes the reference's keys (``img_size``,
    ``channels`` (string), ``learning_rate``, ``beta_1``, ``beta_2``, ``lambda``, ``batch_size``,
    ``seed`` ...) plus two optional ones: ``precision`` ('bf16' default | 'fp32') and ``device``."""

    def __init__(self, config):
        self.config = config
        self.loss_obj = self.loss_object()
        device = int(config.get("device", os.environ.get("LOCAL_RANK", 0)))
        self.ctx = Context(device=device, precision=config.get("precision", "bf16"), seed=int(config.get("seed", 123)))
        self._init_rng = np.random.default_rng(int(config.get("seed", 123)) + 1)
        self._maybe_init_data_parallel()

    def _maybe_init_data_parallel(self):
        try:
            import torch.distributed as dist
        except Exception:  # torch absent: single device
            return
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            rank, world, uid = exchange_unique_id()
            self.ctx.init_data_parallel(rank, world, uid)

    # -- weight init: conv kernels N(0, 0.02), biases 0, BN gamma 1 / beta 0, IN scale N(1, 0.02)
    #    (reference base_gan.py:74,103,132,200; utils.py:17,23).  Generated on the host so that the
    #    CPU oracle and the device share bit-identical parameters.
    def _initialize(self, model: Model, norm_type: str):
        for v in model.trainable_variables:
            if v.name.endswith(".kernel"):
                v.assign(self._init_rng.normal(0.0, 0.02, size=v.shape).astype(np.float32))
            elif v.name.endswith(".gamma"):
                if norm_type == "instancenorm":
                    v.assign(self._init_rng.normal(1.0, 0.02, size=v.shape).astype(np.float32))
                else:
                    v.assign(np.ones(v.shape, dtype=np.float32))
            else:
                v.assign(np.zeros(v.shape, dtype=np.float32))

    def normalize(self, image):
        """Reference base_gan.py:56-61."""
        return (image / 127.5) - 1

    def Generator(self, norm_type="batchnorm", shape: tuple = (None, None, None)):
        """Reference base_gan.py:168-225.  ``shape`` = (H, W, C); None sizes fall back to
        ``config['img_size']`` (CycleGAN builds its generators with (None, None, C))."""
        h = shape[0] or self.config["img_size"]
        w = shape[1] or self.config["img_size"]
        c = int(shape[2] if shape[2] is not None else self.config["channels"])
        handle = C.c_void_p()
        _ffi.check(_ffi.lib().gan_generator_create(self.ctx.handle, _NORMS[norm_type.lower()], h, w, c, C.byref(handle)))
        model = Model(self.ctx, handle, "generator", (h, w), c)
        self._initialize(model, norm_type.lower())
        return model

    def Discriminator(self, norm_type: str = "batchnorm", target: bool = True):
        """Reference base_gan.py:124-166."""
        c = int(self.config["channels"])
        handle = C.c_void_p()
        _ffi.check(_ffi.lib().gan_discriminator_create(self.ctx.handle, _NORMS[norm_type.lower()], c, int(target),
                                                       C.byref(handle)))
        model = Model(self.ctx, handle, "discriminator", (None, None), c, target=target)
        self._initialize(model, norm_type.lower())
        return model

    def loss_object(self):
        """Reference base_gan.py:227-231."""
        return BinaryCrossentropyFromLogits()

    def discriminator_loss(self, real, generated, factor: float = 1.0):
        """Reference base_gan.py:233-245 (host arrays; the train step fuses this on the device)."""
        real_loss = self.loss_obj(np.ones_like(real), real)
        generated_loss = self.loss_obj(np.zeros_like(generated), generated)
        return (real_loss + generated_loss) * factor

    def optimizer(self, learning_rate: float = 2e-4, beta_1: float = 0.5, beta_2: float = 0.999):
        """Reference base_gan.py:247-252."""
        return Adam(learning_rate=learning_rate, beta_1=beta_1, beta_2=beta_2)

    @abstractmethod
    def generator_loss(self, *args, **kwargs):
        return

    @abstractmethod
    def train_step(self, *args, **kwargs):
        return

    @abstractmethod
    def fit(self, *args, **kwargs):
        return

    @abstractmethod
    def predict(self, *args, **kwargs):
        return
