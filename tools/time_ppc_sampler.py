"""Development helper: ppc latent sampler (L = 16, time conditioned) at config-2 size, both tensor-core kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _models
from graspldm_b200 import _lib
dev = torch.device("cuda:0")
m = _models.build("ppc").to(dev)
m.set_inference_timesteps(100)
m.diffusion_model.rng_mode = "fused"
m.diffusion_model.precision = "bf16"
n_obj, G = 64, 20
n = n_obj * G
z = torch.randn(n_obj, 3, 256, device=dev)
x_T = torch.randn(n, 1, 16, device=dev)
for name, flag in (("row-major", -1), ("channel-major", 0)):
    _lib.call("gldm_sampler_tc_set_rows", flag)
    for i in range(2):
        m.diffusion_model.sample(z_cond=z, batch_size=n, x_T=x_T, grasps_per_object=G, seed=i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(5):
        out, _ = m.diffusion_model.sample(z_cond=z, batch_size=n, x_T=x_T, grasps_per_object=G, seed=i)
    b.record(); torch.cuda.synchronize()
    print(f"ppc sampler, {n} samples x 100 DDPM steps, {name}: {a.elapsed_time(b) / 5:.2f} ms  (|x| {out.abs().mean().item():.3f})")
_lib.call("gldm_sampler_tc_set_rows", -1)
