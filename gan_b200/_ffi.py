"""ctypes binding of libgan_b200.so (include/gan_b200.h).

The shared library is built in-tree by ``make -C gan_b200/csrc`` (or ``__graft_entry__.build()``).
There is no fallback: if the library is missing, or there is no sm_100 device when a context is
created, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libgan_b200.so")

GAN_FP32, GAN_BF16 = 0, 1
GAN_NORM_BATCH, GAN_NORM_INSTANCE = 1, 2
GAN_ENGINE_AUTO, GAN_ENGINE_FFMA, GAN_ENGINE_UMMA = -1, 0, 1

_lib = None

c_float_p = C.POINTER(C.c_float)
_vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/gan_b200.h declares.
SIGNATURES = {
    "gan_last_error": (C.c_char_p, []),
    "gan_version": (C.c_int, []),
    "gan_ctx_create": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.POINTER(_vp)]),
    "gan_ctx_destroy": (C.c_int, [_vp]),
    "gan_ctx_sync": (C.c_int, [_vp]),
    "gan_ctx_set_dropout": (C.c_int, [_vp, C.c_int]),
    "gan_ctx_set_rng": (C.c_int, [_vp, C.c_uint64, C.c_uint32]),
    "gan_ctx_get_call_counter": (C.c_int, [_vp, C.POINTER(C.c_uint32)]),
    "gan_ctx_set_engine": (C.c_int, [_vp, C.c_int]),
    "gan_ctx_set_graphs": (C.c_int, [_vp, C.c_int]),
    "gan_ctx_launch_count": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "gan_ctx_set_profile": (C.c_int, [_vp, C.c_int]),
    "gan_ctx_profile_read": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "gan_ctx_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "gan_comm_unique_id": (C.c_int, [_vp]),
    "gan_ctx_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "gan_ctx_set_sample_offset": (C.c_int, [_vp, C.c_int64]),
    "gan_generator_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "gan_discriminator_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "gan_net_destroy": (C.c_int, [_vp]),
    "gan_net_num_tensors": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gan_net_tensor_info": (C.c_int, [_vp, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_int64)]),
    "gan_net_get_tensor": (C.c_int, [_vp, C.c_int, _vp]),
    "gan_net_set_tensor": (C.c_int, [_vp, C.c_int, _vp]),
    "gan_net_get_grad": (C.c_int, [_vp, C.c_int, _vp]),
    "gan_net_num_params": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "gan_net_get_params": (C.c_int, [_vp, _vp]),
    "gan_net_set_params": (C.c_int, [_vp, _vp]),
    "gan_net_get_grads": (C.c_int, [_vp, _vp]),
    "gan_net_debug_tensor": (C.c_int, [_vp, C.c_int, C.c_char_p, _vp, C.c_int64, C.POINTER(C.c_int64)]),
    "gan_generator_forward": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "gan_discriminator_forward": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "gan_adam_create": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(_vp)]),
    "gan_adam_destroy": (C.c_int, [_vp]),
    "gan_adam_get_step": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "gan_adam_set_step": (C.c_int, [_vp, C.c_int64]),
    "gan_adam_get_state": (C.c_int, [_vp, C.c_int, _vp]),
    "gan_adam_set_state": (C.c_int, [_vp, C.c_int, _vp]),
    "gan_adam_set_hyper": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, C.c_double]),
    "gan_pix2pix_train_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_float, C.c_int, _vp]),
    "gan_pix2pix_train_step_ex": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_float, C.c_float, C.c_int, _vp]),
    "gan_cyclegan_train_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_float,
                                          C.c_int, _vp]),
    "gan_ctx_last_losses": (C.c_int, [_vp, _vp, C.c_int]),
    "gan_ctx_prefetch": (C.c_int, [_vp, _vp, _vp, C.c_int64]),
    "gan_preprocess_images": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "gan_ctx_prefetch_images": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, C.c_int64, _vp, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(_vp), C.POINTER(_vp)]),
    "gan_op_conv": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_int]),
}


class ImageXform(C.Structure):
    """``gan_image_xform`` of include/gan_b200.h."""
    _fields_ = [(n, C.c_int) for n in ("src_h", "src_w", "col0", "cols", "pre", "mid", "crop_y", "crop_x", "flip")]


class GanError(RuntimeError):
    pass


def lib():
    """Load libgan_b200.so once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GanError(f"{LIB_PATH} not found: build it with `make -C gan_b200/csrc` "
                       f"(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)            # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def check(status: int):
    if status != 0:
        msg = lib().gan_last_error()
        raise GanError(f"libgan_b200 error {status}: {msg.decode() if msg else ''}")


def ptr_of(x):
    """Address of a host numpy array or of a torch tensor (host or CUDA); NHWC float32 contiguous."""
    if x is None:
        return None
    if hasattr(x, "ptr") and hasattr(x, "data_ptr"):   # input_pipeline.DeviceBatch: borrowed device float32 batch
        return C.c_void_p(x.ptr)
    if hasattr(x, "data_ptr"):                 # torch.Tensor (allocation only; never used for math)
        if str(x.dtype) != "torch.float32" or not x.is_contiguous():
            raise GanError("tensors must be contiguous float32")
        return C.c_void_p(x.data_ptr())
    import numpy as np
    if not isinstance(x, np.ndarray) or x.dtype != np.float32 or not x.flags["C_CONTIGUOUS"]:
        raise GanError("arrays must be C-contiguous float32 numpy arrays")
    return C.c_void_p(x.ctypes.data)
