"""Diagnostic: bf16-path error of the generator forward / first train step vs the fp64 oracle as a
function of batch size (BatchNorm at the 1x1 bottleneck sees n = B samples)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_pix2pix, load_model
from oracle import gan_oracle as O
from gan_b200 import Pix2Pix

def run(prec, B, engine=-1):
    cfg = dict(img_size=256, channels='3', learning_rate=2e-4, beta_1=0.5, beta_2=0.999, generator_loss='l1', seed=123, precision=prec)
    cfg['lambda'] = 100
    m = Pix2Pix(cfg); m.ctx.set_engine(engine)
    g_np, d_np = make_pix2pix(124, 3, None)
    load_model(m.generator, g_np); load_model(m.discriminator, d_np)
    rng = np.random.default_rng(123)
    x = O.synthetic_images(rng, B, 256, 256, 3); y = O.synthetic_images(rng, B, 256, 256, 3)
    gp = O.to_torch(g_np, torch.float64)
    masks = O.generator_keep_masks(123, m.ctx.call_counter(), 0, B, 256)
    out = m.generator(x)
    taps = {}
    ref = O.generator_forward(gp, torch.tensor(x, dtype=torch.float64), "batchnorm", masks, taps=taps).detach().numpy()
    d = out - ref
    line = f"{prec} eng={engine} B={B}: gen_out max_rel={np.abs(d).max()/np.abs(ref).max():.3e} l2_rel={np.linalg.norm(d)/np.linalg.norm(ref):.3e} |"
    for name in ["down1.a", "down4.a", "down6.a", "down7.a", "down8.a", "up1.a", "up2.a", "up4.a", "up7.a"]:
        dev = m.generator.debug_tensor(name); r = taps[name].detach().numpy().reshape(-1)
        line += f" {name}:{np.abs(dev-r).max()/np.abs(r).max():.1e}/{np.linalg.norm(dev-r)/np.linalg.norm(r):.1e}"
    print(line, flush=True)
    m.ctx.close()

for B in (2, 8, 32):
    run("fp32", B)
    run("bf16", B, 0)
    run("bf16", B, -1)
