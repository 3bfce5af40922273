"""Development helper: per-row-job cycle stamps (CTA 0, step 1) of the row-major sampler kernel."""
import os, sys
os.environ["GLDM_TC_ROWS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _models
from graspldm_b200 import _lib
dev = torch.device("cuda:0")
n = 1280
what = sys.argv[1] if len(sys.argv) > 1 else "sampler"      # sampler (fpc, 10 steps) | decoder (one evaluation, L = 16)
m = _models.build("fpc").to(dev)
m.set_inference_timesteps(10)
m.diffusion_model.rng_mode = "fused"
m.diffusion_model.precision = "bf16"
m.vae_model.decoder.precision = "bf16"
z = torch.randn(n // 20, 3, 64, device=dev)
x_T = torch.randn(n, 1, 4, device=dev)
def run(seed):
    if what == "sampler":
        m.diffusion_model.sample(z_cond=z, batch_size=n, x_T=x_T, grasps_per_object=20, seed=seed)
    else:
        m.vae_model.decoder(x_T[:, 0], z, grasps_per_object=20)
run(0)
buf = torch.zeros(1024, dtype=torch.int64, device=dev)
_lib.call("gldm_sampler_tc_set_profile", buf.data_ptr())
run(1)
torch.cuda.synchronize()
_lib.call("gldm_sampler_tc_set_profile", None)
b = buf.cpu().tolist()
print(f"CTA 0: setup {b[1] - b[0]} cycles, whole kernel {b[2] - b[0]} cycles, first job starts {b[64] - b[0]} after entry")
names = []
for st in range(4):
    names += [f"s{st}.rb0.c1", f"s{st}.rb0.c2", f"s{st}.rb1.c1", f"s{st}.rb1.c2", f"s{st}.qkv", f"s{st}.out", f"s{st}.down"]
names += ["fin.c1", "fin.c2"]
prev = None
tot = [0] * 7
print("row-job        wait(umma)  startbar   pass1   exchange   pass2   pass3+commit   gap")
for j, nm in enumerate(names):
    t = b[64 + 8 * j: 64 + 8 * j + 8]
    if t[7]: print(f"      (TMEM load + wait: {t[7] - t[2]} cycles)")
    gap = (t[0] - prev) if prev is not None else 0
    if t[3] == 0:
        row = [t[1] - t[0], t[2] - t[1], 0, 0, t[6] - t[2], 0, gap]
    else:
        row = [t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5], gap]
    tot = [x + y for x, y in zip(tot, row)]
    print(f"{nm:14s} " + " ".join(f"{x:9d}" for x in row))
    prev = t[6]
print("total          " + " ".join(f"{x:9d}" for x in tot), "  first..last", b[64 + 8 * (len(names) - 1) + 6] - b[64])

