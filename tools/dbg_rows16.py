"""Development helper: L = 16 row-major kernel against the fp32 path, error pattern per (sample, position)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import _models
np.set_printoptions(linewidth=220, precision=3, suppress=True)
dev = torch.device("cuda:0")
m = _models.build("ppc").to(dev)
gen = torch.Generator().manual_seed(5)
n, gpo = int(sys.argv[1]) if len(sys.argv) > 1 else 12, 4
x = torch.randn(n, 1, 16, generator=gen).to(dev)
t = torch.randint(0, 1000, (n,), generator=gen).to(dev)
z = torch.randn(n, 3, 256, generator=gen).to(dev)
net = m.diffusion_model.model
a = net(x, time=t, z_cond=z).cpu().numpy().reshape(n, 16)
b = net(x, time=t, z_cond=z, precision="bf16").cpu().numpy().reshape(n, 16)
print("max|eps|", np.abs(a).max(), "max err", np.abs(a - b).max())
print("err per (sample, position):")
print(np.abs(a - b))
# decoder
dec = m.vae_model.decoder
zh = torch.randn(n, 16 if hasattr(dec, "in_features") and dec.in_features == 16 else dec.in_features, generator=gen).to(dev)
zc = torch.randn((n + gpo - 1) // gpo, 3, 256, generator=gen).to(dev)
t32, l32 = dec(zh, zc[: n // gpo] if n % gpo == 0 else zc.repeat_interleave(gpo, 0)[:n], grasps_per_object=gpo if n % gpo == 0 else 1)
dec.precision = "bf16"
t16, l16 = dec(zh, zc[: n // gpo] if n % gpo == 0 else zc.repeat_interleave(gpo, 0)[:n], grasps_per_object=gpo if n % gpo == 0 else 1)
print("decoder tmrp err", (t32 - t16).abs().max().item(), "logit err", (l32 - l16).abs().max().item(), "max|tmrp|", t32.abs().max().item())
print((t32 - t16).abs().cpu().numpy())
