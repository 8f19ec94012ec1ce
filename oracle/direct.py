"""Direct-definition NumPy restatement of the layer arithmetic (tiny shapes only).

TEST INFRASTRUCTURE ONLY (see oracle/gan_oracle.py header; parity unpinned against TF).
These loops follow the index formulas of SURVEY.md Appendix A term by term and exist to pin the
padding / no-kernel-flip / kernel-layout conventions of ``gan_oracle`` (which goes through
torch's conv routines) with an independent formulation, and to pin the tap tables that the CUDA
kernels use (the parity-class decomposition of the transposed convolution).
"""
import numpy as np


def conv2d_s2_same(x, w):
    """y[n,oh,ow,co] = sum_{kh,kw,ci} xpad[n,2oh+kh,2ow+kw,ci] w[kh,kw,ci,co], pad (1,1) (App. A.2;
    base_gan.py:78-79)."""
    n, h, wd, ci = x.shape
    co = w.shape[3]
    xp = np.zeros((n, h + 2, wd + 2, ci), dtype=np.float64)
    xp[:, 1:-1, 1:-1] = x
    y = np.zeros((n, h // 2, wd // 2, co), dtype=np.float64)
    for oh in range(h // 2):
        for ow in range(wd // 2):
            patch = xp[:, 2 * oh:2 * oh + 4, 2 * ow:2 * ow + 4, :]          # n,kh,kw,ci
            y[:, oh, ow, :] = np.einsum("nhwc,hwco->no", patch, w)
    return y


def conv2d_s1_pad1(x, w, bias=None):
    """ZeroPadding2D(1) then 4x4 stride-1 'valid' conv (App. A.4; base_gan.py:145-148,157-161)."""
    n, h, wd, ci = x.shape
    co = w.shape[3]
    xp = np.zeros((n, h + 2, wd + 2, ci), dtype=np.float64)
    xp[:, 1:-1, 1:-1] = x
    y = np.zeros((n, h - 1, wd - 1, co), dtype=np.float64)
    for oh in range(h - 1):
        for ow in range(wd - 1):
            y[:, oh, ow, :] = np.einsum("nhwc,hwco->no", xp[:, oh:oh + 4, ow:ow + 4, :], w)
    if bias is not None:
        y += bias
    return y


def conv2d_transpose_s2_same(x, f, bias=None):
    """y[n,oh,ow,co] = sum over (h,kh): oh=2h+kh-1, (w,kw): ow=2w+kw-1 of x[n,h,w,ci] f[kh,kw,co,ci]
    (App. A.3; base_gan.py:107-110). Scatter form."""
    n, h, wd, ci = x.shape
    co = f.shape[2]
    y = np.zeros((n, 2 * h, 2 * wd, co), dtype=np.float64)
    for ih in range(h):
        for iw in range(wd):
            for kh in range(4):
                for kw in range(4):
                    oh, ow = 2 * ih + kh - 1, 2 * iw + kw - 1
                    if 0 <= oh < 2 * h and 0 <= ow < 2 * wd:
                        y[:, oh, ow, :] += x[:, ih, iw, :] @ f[kh, kw].T
    if bias is not None:
        y += bias
    return y


# Parity-class tap table of the transposed convolution (App. A.3): for output parity a,
# taps are (kh, dh) with input row = i + dh for output row 2i + a.
CONVT_TAPS = {0: ((1, 0), (3, -1)), 1: ((0, 1), (2, 0))}


def conv2d_transpose_s2_same_gather(x, f):
    """Gather (parity-class) form of the transposed convolution, the form the GPU kernels use:
    four output classes (a,b), each a 2x2 stride-1 correlation with K = 4*Cin."""
    n, h, wd, ci = x.shape
    co = f.shape[2]
    y = np.zeros((n, 2 * h, 2 * wd, co), dtype=np.float64)
    for a in (0, 1):
        for b in (0, 1):
            for i in range(h):
                for j in range(wd):
                    acc = np.zeros((n, co))
                    for kh, dh in CONVT_TAPS[a]:
                        for kw, dw in CONVT_TAPS[b]:
                            ih, iw = i + dh, j + dw
                            if 0 <= ih < h and 0 <= iw < wd:
                                acc += x[:, ih, iw, :] @ f[kh, kw].T
                    y[:, 2 * i + a, 2 * j + b, :] = acc
    return y


def batch_norm_train(x, gamma, beta, eps=1e-3):
    mean = x.mean(axis=(0, 1, 2))
    var = ((x - mean) ** 2).mean(axis=(0, 1, 2))
    return gamma * (x - mean) / np.sqrt(var + eps) + beta


def instance_norm(x, scale, offset, eps=1e-5):
    mean = x.mean(axis=(1, 2), keepdims=True)
    var = ((x - mean) ** 2).mean(axis=(1, 2), keepdims=True)
    return scale * (x - mean) / np.sqrt(var + eps) + offset


def bce_from_logits(x, z):
    return float(np.mean(np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))))


def keras_adam_step(theta, g, m, v, t, lr=2e-4, b1=0.5, b2=0.999, eps=1e-7):
    """One Keras-Adam update (App. A.11), returns (theta, m, v)."""
    m = m + (g - m) * (1 - b1)
    v = v + (g * g - v) * (1 - b2)
    alpha = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return theta - alpha * m / (np.sqrt(v) + eps), m, v
