"""gan_b200 — B200-native (sm_100a) implementation of the kingjosephm/GAN train step.

Host classes mirror the reference (`GAN`, `Pix2Pix`, `CycleGAN`); all arithmetic runs in
`csrc/libgan_b200.so` through the C-ABI declared in `include/gan_b200.h`.
"""
from .base_gan import GAN, Context, Model, Adam, LossValue, nccl_unique_id, exchange_unique_id, shard_bounds  # noqa: F401
from .pix2pix import Pix2Pix  # noqa: F401
from .cycle_gan import CycleGAN  # noqa: F401
from . import _ffi, checkpoint, input_pipeline  # noqa: F401
from .checkpoint import Checkpoint, CheckpointManager, latest_checkpoint  # noqa: F401

__all__ = ["GAN", "Pix2Pix", "CycleGAN", "Context", "Model", "Adam", "LossValue", "Checkpoint", "CheckpointManager",
           "latest_checkpoint"]
