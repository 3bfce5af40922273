"""Development helper: TMEM dumps of the L = 16 row-major kernel (CTA 0) against torch intermediates, job by job."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.nn.functional as F
import _models
from graspldm_b200 import _lib
from oracle import model_torch as M
np.set_printoptions(linewidth=220, precision=3, suppress=True)
dev = torch.device("cuda:0")
m = _models.build("ppc").to(dev)
net = m.diffusion_model.model
sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
gen = torch.Generator().manual_seed(5)
n = 8
x = torch.randn(n, 1, 16, generator=gen)
t = torch.randint(0, 1000, (n,), generator=gen)
z = torch.randn(n, 3, 256, generator=gen)
p = ""
# ---- torch intermediates of stage 0 (pre-epilogue accumulators = conv outputs without bias)
emb = M.time_embedding(sd, p, t)
inp = F.silu(F.linear(z, sd[p + "input_emb_layers.0.weight"], sd[p + "input_emb_layers.0.bias"]))
emb = emb.unsqueeze(-2).repeat(1, inp.shape[1], 1) + inp
def ws(w):
    mean = w.mean(dim=(1, 2), keepdim=True); var = w.var(dim=(1, 2), unbiased=False, keepdim=True)
    return (w - mean) * (var + 1e-5).rsqrt()
accs, names = [], []
h = F.conv1d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=3)
res0 = h
def rb(pfx, xin):
    e = F.linear(F.silu(emb), sd[pfx + "mlp.1.weight"], sd[pfx + "mlp.1.bias"]).transpose(1, 2)
    scale, shift = e.chunk(2, dim=1)
    a1 = F.conv1d(xin, ws(sd[pfx + "block1.proj.weight"]), None, padding=1); accs.append(a1); names.append(pfx + "c1")
    y = F.group_norm(a1 + sd[pfx + "block1.proj.bias"].view(1, -1, 1), 4, sd[pfx + "block1.norm.weight"], sd[pfx + "block1.norm.bias"], 1e-5)
    y = (y.unsqueeze(-1) * (scale.unsqueeze(-2) + 1) + shift.unsqueeze(-2)).sum(-1)
    y = F.silu(y)
    a2 = F.conv1d(y, ws(sd[pfx + "block2.proj.weight"]), None, padding=1); accs.append(a2); names.append(pfx + "c2")
    y2 = F.silu(F.group_norm(a2 + sd[pfx + "block2.proj.bias"].view(1, -1, 1), 4, sd[pfx + "block2.norm.weight"], sd[pfx + "block2.norm.bias"], 1e-5))
    return y2 + xin
h = rb("blocks.0.0.", h)
h = rb("blocks.0.1.", h)
xn = M._chan_layernorm(h, sd["blocks.0.2.fn.norm.g"])
qkv = F.conv1d(xn, sd["blocks.0.2.fn.fn.to_qkv.weight"]); accs.append(qkv); names.append("qkv")
B_, C_, n_ = h.shape
q, k, v = qkv.view(B_, 3, 4, 32, n_).unbind(1)
q = q.softmax(dim=-2) * 32 ** -0.5
k = k.softmax(dim=-1)
ctx = torch.einsum("bhdn,bhen->bhde", k, v)
ao = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(B_, 128, n_)
o = F.conv1d(ao, sd["blocks.0.2.fn.fn.to_out.0.weight"], None); accs.append(o); names.append("out")
o = M._chan_layernorm(o + sd["blocks.0.2.fn.fn.to_out.0.bias"].view(1, -1, 1), sd["blocks.0.2.fn.fn.to_out.1.g"]) + h
d = F.conv1d(o, sd["blocks.0.3.weight"], None, padding=1); accs.append(d); names.append("down")

buf = torch.zeros(16 + 128 * 512 // 2 + 64, dtype=torch.int64, device=dev)
xd, td, zd = x.to(dev), t.to(dev), z.to(dev)
for j, (nm, want) in enumerate(zip(names, accs)):
    buf.zero_()
    buf[8] = j + 1
    _lib.call("gldm_sampler_tc_set_profile", buf.data_ptr())
    net(xd, time=td, z_cond=zd, precision="bf16")
    torch.cuda.synchronize()
    _lib.call("gldm_sampler_tc_set_profile", None)
    img = buf[16:16 + 128 * 512 // 2].view(torch.float32).view(128, 512).cpu()
    C = want.shape[1]
    got = img[:, :C].view(16, 8, C).permute(1, 2, 0)[:n]       # row = pos * 8 + s -> [s][c][pos]
    err = (got - want).abs()
    print(f"job {j} {nm:18s} C={C:3d} max|want| {want.abs().max():.3f} max err {err.max():.3e}  per-position max:", err.amax(dim=(0, 1)).numpy())
    if j == 0:
        r = img[:, 384:384 + 16].view(16, 8, 16).permute(1, 2, 0)[:n]
        print("   residual (init conv) err", (r - res0).abs().max().item())
    if err.max() > 0.1:
        print("   per-channel max:", err.amax(dim=(0, 2)).numpy())
        print("   per-sample max:", err.amax(dim=(1, 2)).numpy())
