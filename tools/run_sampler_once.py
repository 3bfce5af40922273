"""Development helper: one bf16 sampler launch (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _models
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1280
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
sets = int(sys.argv[4]) if len(sys.argv) > 4 else 0
if sets:
    from graspldm_b200 import _lib
    _lib.call("gldm_sampler_tc_set_sets", sets)
dev = torch.device("cuda:0")
model = _models.build("fpc").to(dev)
model.set_inference_timesteps(steps)
model.diffusion_model.rng_mode = "fused"
model.diffusion_model.precision = prec
z = torch.randn(n // 20, 3, 64, device=dev)
x_T = torch.randn(n, 1, 4, device=dev)
for i in range(2):
    out, _ = model.diffusion_model.sample(z_cond=z, batch_size=n, x_T=x_T, grasps_per_object=20, seed=i)
torch.cuda.synchronize()
print("ok", out.abs().mean().item())
