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


# ---------------------------------------------------------------------------------------------------------------
# Restatements of the operand layouts of the first-layer kernels (gan_b200/csrc/conv_first.cu) and of the per-tap data
# gradient of the discriminator's first layer (engine.cu:discriminator_backward, elem.cu:k_col2im_grad).  They pin the
# index maps on the CPU: k = tap*4 + channel slot, tap = kh*4 + kw, and the parity tap table of the gather.
# ---------------------------------------------------------------------------------------------------------------
def slot4_rows(x):
    """rows[n, oh, ow, (kh*4+kw)*4 + c] = xpad[n, 2oh+kh, 2ow+kw, c] for c < C <= 4, zero in the slots c >= C: the
    128-byte operand row that k_conv_first_fwd / k_conv_first_wgrad assemble in shared memory (conv_first.cu:build_row)."""
    n, h, w, c = x.shape
    assert c <= 4
    xp = np.zeros((n, h + 2, w + 2, c), dtype=np.float64)
    xp[:, 1:-1, 1:-1] = x
    rows = np.zeros((n, h // 2, w // 2, 64), dtype=np.float64)
    for kh in range(4):
        for kw in range(4):
            rows[..., (kh * 4 + kw) * 4:(kh * 4 + kw) * 4 + c] = xp[:, kh:kh + h:2, kw:kw + w:2, :]
    return rows


def slot4_weights(w, c0, c):
    """B[co, (kh*4+kw)*4 + s] = w[kh, kw, c0 + s, co]: the first layer's weight pack for the source image whose channels
    start at c0 (engine.cu:pack_weights, wp_im2col)."""
    co = w.shape[3]
    b = np.zeros((co, 64), dtype=np.float64)
    for t in range(16):
        b[:, t * 4:t * 4 + c] = w[t // 4, t % 4, c0:c0 + c, :].T
    return b


def conv2d_s2_same_via_rows(sources, w):
    """Conv2D 4x4 s2 'same' of concatenate(sources, -1) as the sum over sources of rows @ B^T (the first-layer GEMM)."""
    c = sources[0].shape[3]
    return sum(slot4_rows(x) @ slot4_weights(w, i * c, c).T for i, x in enumerate(sources))


def conv2d_s2_same_wgrad_via_rows(sources, dz):
    """dW[kh, kw, i*C + s, co] = sum over output pixels of rows_i[.., (kh*4+kw)*4+s] * dz[.., co] (k_conv_first_wgrad +
    k_wgrad_reduce's im2col mapping: master offset (k>>2)*Cin*Cout + (source*C + (k&3))*Cout + co)."""
    c = sources[0].shape[3]
    co = dz.shape[3]
    dw = np.zeros((4, 4, c * len(sources), co), dtype=np.float64)
    for i, x in enumerate(sources):
        g = np.einsum("nhwk,nhwo->ko", slot4_rows(x), dz)                # [64][co]
        for k in range(64):
            if (k & 3) < c:
                dw[(k >> 2) // 4, (k >> 2) % 4, i * c + (k & 3), :] = g[k]
    return dw


def conv2d_s2_same_dgrad_via_cols(dz, w, c0, c):
    """d/dx[.., c0:c0+c] of Conv2D 4x4 s2 'same': cols[m, tap*4+s] = dz[m, :] . w[kh, kw, c0+s, :] (ONE 1x1 GEMM), then the
    parity gather of k_col2im_grad: dx[2i+a, 2j+b] = sum over CONVT_TAPS[a] x CONVT_TAPS[b] of cols[i+dh, j+dw][tap]."""
    n, ho, wo, _ = dz.shape
    cols = np.zeros((n, ho, wo, 64), dtype=np.float64)
    for t in range(16):
        cols[..., t * 4:t * 4 + c] = dz @ w[t // 4, t % 4, c0:c0 + c, :].T
    dx = np.zeros((n, 2 * ho, 2 * wo, c), dtype=np.float64)
    for a in (0, 1):
        for b in (0, 1):
            for kh, dh in CONVT_TAPS[a]:
                for kw, dw in CONVT_TAPS[b]:
                    for i in range(ho):
                        for j in range(wo):
                            ih, iw = i + dh, j + dw
                            if 0 <= ih < ho and 0 <= iw < wo:
                                dx[:, 2 * i + a, 2 * j + b, :] += cols[:, ih, iw, (kh * 4 + kw) * 4:(kh * 4 + kw) * 4 + c]
    return dx
