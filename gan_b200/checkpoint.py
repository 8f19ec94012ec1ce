"""Weight / optimizer import-export in Keras variable order.

Stands in for the reference's ``tf.train.Checkpoint(generator_optimizer=..., generator=..., ...)``
and ``tf.train.CheckpointManager(checkpoint, dir, max_to_keep)`` (pix2pix.py:399-422,308-317,
cycle_gan.py:436-463,342-350).  The TF tensor-bundle format needs TensorFlow; this is an own format,
one ``.npz`` per save:

    <object>/<variable name>            every tensor of a Model in ``model.variables`` order
                                        (trainable ones first, Keras order App. A.8, then the BatchNorm
                                        moving_mean / moving_variance pairs)
    <object>/iterations                 Adam step counter (int64)
    <object>/m/<variable name>, <object>/v/<variable name>      Adam slots, one per trainable variable
    <object>/hyper                      [learning_rate, beta_1, beta_2, epsilon]
    __meta__                            JSON: format version, object kinds, dropout call counter

so a converter from a real TF run only has to rename ``layer/kernel:0``-style keys.  Everything moves
through the C-ABI getters/setters (gan_net_get/set_tensor, gan_adam_get/set_state, get/set_step).
"""
from __future__ import annotations

import ctypes as C
import glob
import json
import os
import re

import numpy as np

from . import _ffi
from .base_gan import Adam, Model

FORMAT_VERSION = 1


class _Status:
    """Return value of ``restore``; the reference calls ``.expect_partial()`` on it (pix2pix.py:411)."""

    def __init__(self, missing, unused):
        self.missing, self.unused = missing, unused

    def expect_partial(self):
        return self

    def assert_consumed(self):
        if self.missing or self.unused:
            raise AssertionError(f"checkpoint not fully matched: missing {self.missing[:5]}, unused {self.unused[:5]}")
        return self


class Checkpoint:
    def __init__(self, **objects):
        for k, v in objects.items():
            if not isinstance(v, (Model, Adam)):
                raise TypeError(f"{k}: only Model and Adam objects can be checkpointed")
        self._objects = dict(objects)
        self.ctx = None
        for v in objects.values():
            if isinstance(v, Model):
                self.ctx = v._ctx

    # -- save ------------------------------------------------------------------------------------
    def _collect(self):
        out, kinds = {}, {}
        for name, obj in self._objects.items():
            if isinstance(obj, Model):
                kinds[name] = obj.kind
                for v in obj.variables:
                    out[f"{name}/{v.name}"] = v.numpy()
            else:
                kinds[name] = "adam"
                out[f"{name}/hyper"] = np.array([obj.learning_rate, obj.beta_1, obj.beta_2, obj.epsilon], dtype=np.float64)
                out[f"{name}/iterations"] = np.array(obj.iterations, dtype=np.int64)
                if obj._h is not None:                      # slots exist once the optimizer has been bound
                    m, v = obj.get_state("m"), obj.get_state("v")
                    off = 0
                    for var in obj._model.trainable_variables:
                        n = int(np.prod(var.shape))
                        out[f"{name}/m/{var.name}"] = m[off:off + n].reshape(var.shape)
                        out[f"{name}/v/{var.name}"] = v[off:off + n].reshape(var.shape)
                        off += n
        meta = {"format": FORMAT_VERSION, "objects": kinds,
                "call_counter": self.ctx.call_counter() if self.ctx is not None else 0}
        out["__meta__"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        return out

    def write(self, file_prefix: str) -> str:
        path = file_prefix + ".npz"
        if self.ctx is not None and getattr(self.ctx, "rank", 0) != 0:
            return path                                     # data parallel: replicas are identical, rank 0 writes
        d = os.path.dirname(os.path.abspath(path))
        os.makedirs(d, exist_ok=True)
        # np.savez appends ".npz" to names without it; the leading dot keeps the glob "ckpt-*.npz" from matching
        tmp = os.path.join(d, f".{os.path.basename(file_prefix)}.{os.getpid()}.tmp.npz")
        np.savez(tmp, **self._collect())
        os.replace(tmp, path)                               # a crash never leaves a half-written checkpoint
        return path

    save = write

    # -- restore ---------------------------------------------------------------------------------
    def restore(self, path: str, models_for_optimizers: dict | None = None, restore_rng: bool = True) -> _Status:
        """Load ``path`` (as returned by ``write`` / ``latest_checkpoint``).  Adam slots need the optimizer
        to be bound to its model: pass ``{'generator_optimizer': model, ...}`` or bind beforehand
        (``train_step`` binds at first use); an optimizer that is still unbound keeps its slots and applies them
        when it is bound (TF's deferred restoration), which is what the reference's predict flow relies on
        (pix2pix.py:400-411).  Saved hyper-parameters replace the optimizer's.  The dropout call counter of the
        saved run is restored unless ``restore_rng=False``.  Keys present on only one side are reported, not fatal."""
        if path is None:
            raise ValueError("no checkpoint to restore (latest_checkpoint returned None)")
        data = np.load(path)
        keys = set(data.files) - {"__meta__"}
        used, missing = set(), []
        for name, obj in self._objects.items():
            if isinstance(obj, Model):
                for v in obj.variables:
                    k = f"{name}/{v.name}"
                    if k in keys:
                        v.assign(data[k]); used.add(k)
                    else:
                        missing.append(k)
            else:
                if f"{name}/iterations" not in keys:
                    missing.append(f"{name}/iterations")
                    continue
                model = (models_for_optimizers or {}).get(name, obj._model)
                # tf.train.Checkpoint restores the saved hyper-parameters silently (they are checkpointed variables)
                lr, b1, b2, eps = (float(x) for x in data[f"{name}/hyper"])
                obj.set_hyper(lr, b1, b2, eps)
                used.update({f"{name}/hyper", f"{name}/iterations"})
                slot_keys = {k for k in keys if k.startswith(f"{name}/m/") or k.startswith(f"{name}/v/")}
                if model is None:
                    # The reference restores into a fresh model before any train_step (pix2pix.py:400-411): the
                    # optimizer has no slots yet.  As TF does (deferred restoration), keep the values and apply
                    # them when the optimizer is bound; until then they count as consumed-later, not as errors.
                    obj._deferred = {"iterations": int(data[f"{name}/iterations"]),
                                     "slots": {k[len(name) + 1:]: np.asarray(data[k], np.float32) for k in slot_keys}}
                    used.update(slot_keys)
                    continue
                obj.bind(model)
                state = {"iterations": int(data[f"{name}/iterations"]), "slots": {}}
                for which in ("m", "v"):
                    for var in model.trainable_variables:
                        k = f"{name}/{which}/{var.name}"
                        if k in keys:
                            if data[k].shape != var.shape:
                                raise ValueError(f"{k}: shape {data[k].shape} != {var.shape}")
                            state["slots"][f"{which}/{var.name}"] = np.asarray(data[k], np.float32); used.add(k)
                        else:
                            missing.append(k)
                obj._apply_state(state)
        if "__meta__" in data.files and self.ctx is not None:
            meta = json.loads(bytes(data["__meta__"]).decode())
            self.last_meta = meta
            # the dropout stream continues where the saved run stopped (same seed, saved call counter)
            if restore_rng and "call_counter" in meta:
                self.ctx.set_rng(self.ctx.seed, int(meta["call_counter"]))
        return _Status(missing, sorted(keys - used))


def latest_checkpoint(directory: str):
    """Role of ``tf.train.latest_checkpoint`` (pix2pix.py:411): newest ``ckpt-<n>.npz`` or None."""
    best, best_n = None, -1
    for f in glob.glob(os.path.join(directory, "ckpt-*.npz")):
        m = re.search(r"ckpt-(\d+)\.npz$", f)
        if m and int(m.group(1)) > best_n:
            best, best_n = f, int(m.group(1))
    return best


class CheckpointManager:
    """``tf.train.CheckpointManager(checkpoint, directory, max_to_keep)``: numbered saves, oldest deleted."""

    def __init__(self, checkpoint: Checkpoint, directory: str, max_to_keep: int = 1):
        self.checkpoint, self.directory, self.max_to_keep = checkpoint, directory, max_to_keep
        latest = latest_checkpoint(directory)
        self._n = int(re.search(r"ckpt-(\d+)\.npz$", latest).group(1)) if latest else 0

    @property
    def latest_checkpoint(self):
        return latest_checkpoint(self.directory)

    @property
    def checkpoints(self):
        fs = [f for f in glob.glob(os.path.join(self.directory, "ckpt-*.npz")) if re.search(r"ckpt-(\d+)\.npz$", f)]
        return sorted(fs, key=lambda f: int(re.search(r"ckpt-(\d+)\.npz$", f).group(1)))

    def save(self) -> str:
        self._n += 1
        path = self.checkpoint.write(os.path.join(self.directory, f"ckpt-{self._n}"))
        if self.max_to_keep:
            if getattr(self.checkpoint.ctx, "rank", 0) == 0:
                for old in self.checkpoints[:-self.max_to_keep]:
                    os.remove(old)
        return path
