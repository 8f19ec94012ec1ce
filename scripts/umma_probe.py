"""Diagnostic (not a test): run every tcgen05 case and print error statistics instead of asserting.
Usage on the GPU box: python scripts/umma_probe.py [fwd|dgrad|wgrad ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_conv, oracle_conv_grads, rel_err, bf16_round  # noqa: E402
from gan_b200 import Context  # noqa: E402

roles = sys.argv[1:] or ["fwd", "dgrad", "wgrad"]
CASES = [(0, 2, 16, 16, 64, 128), (0, 1, 32, 32, 128, 64), (0, 4, 4, 4, 64, 64), (1, 2, 8, 8, 64, 128),
         (1, 2, 9, 11, 64, 64), (2, 2, 8, 8, 128, 64), (2, 5, 2, 2, 64, 64), (0, 8, 64, 64, 128, 256),
         (1, 8, 32, 32, 256, 512), (2, 8, 32, 32, 512, 128)]
ctx = Context(0, "bf16", 1)
for kind, b, h, w, cin, cout in CASES:
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, size=(b, h, w, cin)).astype(np.float32)
    wt = rng.normal(0, 0.05, size=((4, 4, cout, cin) if kind == 2 else (4, 4, cin, cout))).astype(np.float32)
    ho, wo = {0: (h // 2, w // 2), 1: (h - 1, w - 1), 2: (2 * h, 2 * w)}[kind]
    dy = rng.normal(0, 1, size=(b, ho, wo, cout)).astype(np.float32)
    xq, wq, dyq = bf16_round(x), bf16_round(wt), bf16_round(dy)
    y_ref = oracle_conv(kind, xq, wq).numpy()
    dx_ref, dw_ref = oracle_conv_grads(kind, xq, wq, dyq)
    for role in roles:
        try:
            t = time.time()
            if role == "fwd":
                out = ctx.op_conv(kind, 0, x, wt, b, h, w, cin, cout, engine=1); ref = y_ref
            elif role == "dgrad":
                out = ctx.op_conv(kind, 1, dy, wt, b, h, w, cin, cout, engine=1); ref = dx_ref
            else:
                out = ctx.op_conv(kind, 2, x, dy, b, h, w, cin, cout, engine=1); ref = dw_ref
            e = rel_err(out, ref)
            bad = np.abs(out - ref) > 2e-2 * np.abs(ref).max()
            print(f"kind={kind} B={b} H={h} W={w} Cin={cin} Cout={cout} {role}: rel_err={e:.3e} "
                  f"bad_frac={bad.mean():.4f} nan={np.isnan(out).sum()} |out|max={np.abs(out).max():.3e} "
                  f"|ref|max={np.abs(ref).max():.3e} t={time.time() - t:.2f}s", flush=True)
            if bad.mean() > 0 and bad.mean() < 1:
                idx = np.argwhere(bad)
                print("   first bad idx:", idx[:4].tolist(), " last bad idx:", idx[-2:].tolist(), flush=True)
        except Exception as ex:  # noqa: BLE001
            print(f"kind={kind} {role}: EXCEPTION {ex}", flush=True)
print("probe done", flush=True)
