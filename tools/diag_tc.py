"""Development diagnostics: per-iteration timings of the generation path sections (bf16 sampler)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _data, _models
from graspldm_b200 import engine
from graspldm_b200.inference import InferenceLDM, default_metas

import gc
if os.environ.get("DIAG_NOGC"):
    gc.disable()
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n_obj = int(sys.argv[2]) if len(sys.argv) > 2 else 64
G = int(sys.argv[3]) if len(sys.argv) > 3 else 20
if os.environ.get("GLDM_TC_SETS"):
    from graspldm_b200 import _lib as _l
    _l.call("gldm_sampler_tc_set_sets", int(os.environ["GLDM_TC_SETS"]))
model = _models.build("fpc").to(dev)
model.set_inference_timesteps(100)
model.diffusion_model.rng_mode = "fused"
model.diffusion_model.precision = prec
model.vae_model.encoder.pc_encoder.precision = prec
model.vae_model.decoder.precision = prec
inf = InferenceLDM(model, device=dev)
pcs = _data.synthetic_clouds(n_obj, 1024, seed=1234, dist="S")
pcs_dev = pcs.to(dev)
pcs_pin = pcs.pin_memory()
metas = default_metas(n_obj)
metas_dev = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in metas.items()}

def ev():
    return torch.cuda.Event(enable_timing=True)

def timeit(fn, n=int(os.environ.get("DIAG_N", "6")), label=""):
    ts, hs = [], []
    for i in range(n):
        torch.cuda.synchronize()
        a, b = ev(), ev()
        t0 = time.perf_counter()
        a.record(); fn(i); b.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b)); hs.append((t1 - t0) * 1e3)
    print(f"{label:34s} gpu ms: " + " ".join(f"{t:7.2f}" for t in ts) + "   host ms: " + " ".join(f"{t:6.2f}" for t in hs), flush=True)

z = model.vae_model.encode_pc(pcs_dev)
x_T = torch.randn(n_obj * G, 1, 4).to(dev)
timeit(lambda i: model.vae_model.encode_pc(pcs_dev), label="encoder")
timeit(lambda i: model.diffusion_model.sample(z_cond=z, batch_size=n_obj * G, x_T=x_T, grasps_per_object=G, seed=i), label=f"sampler {prec}")
lat = x_T[:, 0]
timeit(lambda i: model.vae_model.decoder(lat, z, grasps_per_object=G), label="decoder")
timeit(lambda i: inf.generate_grasps(pcs_dev, metas_dev, num_grasps=G, seed=i), label="generate resident")
timeit(lambda i: inf.generate_grasps(pcs_pin, metas, num_grasps=G, seed=i), label="generate e2e (host inputs)")
engine.SECTIONS.enabled = True
timeit(lambda i: inf.generate_grasps(pcs_dev, metas_dev, num_grasps=G, seed=i), label="generate resident + sections")
print({k: [round(x, 2) for x in v] for k, v in engine.SECTIONS.collect().items()})
if prec == "bf16":
    from graspldm_b200 import _lib
    buf = torch.zeros(512, dtype=torch.int64, device=dev)
    _lib.call("gldm_sampler_tc_set_profile", buf.data_ptr())
    model.diffusion_model.sample(z_cond=z, batch_size=n_obj * G, x_T=x_T, grasps_per_object=G, seed=1)
    torch.cuda.synchronize()
    _lib.call("gldm_sampler_tc_set_profile", None)
    b = buf.cpu().tolist()
    names = []
    for st in range(4):
        names += [f"s{st}.rb0.c1", f"s{st}.rb0.c2", f"s{st}.rb1.c1", f"s{st}.rb1.c2", f"s{st}.qkv", f"s{st}.out", f"s{st}.down"]
    names += ["fin.c1", "fin.c2"]
    print(f"step cycles: {b[321] - b[320]}")
    prev_end = b[320]
    print("job         epilogue  drv:refill+bwait  mma-issue  commit+refill  acc-wait  fullwait0 fullwait1")
    tot = [0] * 7
    for j, nm in enumerate(names):
        w0, w1, d0, d1, d2, d3, wf, we = b[8 * j:8 * j + 8]
        row = [d0 - prev_end, d1 - d0, d2 - d1, d3 - d2, w1 - w0, wf, we]
        tot = [a_ + b_ for a_, b_ in zip(tot, row)]
        print(f"{nm:10s} " + " ".join(f"{x:9d}" for x in row))
        prev_end = w1
    print("total      " + " ".join(f"{x:9d}" for x in tot), " tail", b[321] - prev_end)
