"""Development helper: a few encoder passes (for ncu launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _data, _models
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dev = torch.device("cuda:0")
model = _models.build("fpc").to(dev)
enc = model.vae_model.encoder.pc_encoder
enc.precision = prec
xyz = _data.synthetic_clouds(B, seed=1).to(dev)
for i in range(reps):
    z = model.vae_model.encode_pc(xyz)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
if os.environ.get("GLDM_PROFILE_RANGE"):
    torch.cuda.profiler.start()          # ncu --profile-from-start off: only the timed passes are captured
for i in range(iters):
    z = model.vae_model.encode_pc(xyz)
ev[1].record(); torch.cuda.synchronize()
if os.environ.get("GLDM_PROFILE_RANGE"):
    torch.cuda.profiler.stop()
print("encoder ms", ev[0].elapsed_time(ev[1]) / max(iters, 1), "z", z.abs().mean().item())
