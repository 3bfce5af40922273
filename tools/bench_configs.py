#!/usr/bin/env python
"""Secondary measurements on the other BASELINE.json configurations (not bench lines).
Prints a markdown table; run on a B200:  python tools/bench_configs.py > profiles/rNN_configs.md"""
import gc
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import _data  # noqa: E402
import _models  # noqa: E402
from graspldm_b200.inference import InferenceLDM, default_metas  # noqa: E402

dev = torch.device("cuda:0")
if os.environ.get("GLDM_TC_SETS"):
    from graspldm_b200 import _lib as _l
    _l.call("gldm_sampler_tc_set_sets", int(os.environ["GLDM_TC_SETS"]))
PEAK = 1392.3e12   # measured sustained bf16 (MEASURED_PEAKS.json)


def set_precision(m, prec):
    m.diffusion_model.precision = prec
    m.vae_model.encoder.pc_encoder.precision = prec
    m.vae_model.decoder.precision = prec


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    gc.collect(); gc.disable()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    gc.enable()
    return a.elapsed_time(b) / iters


def ldm(name, sched, steps, n_obj, grasps, prec="bf16"):
    m = _models.build(name, scheduler=sched).to(dev)
    m.set_inference_timesteps(steps)
    m.diffusion_model.rng_mode = "fused"
    set_precision(m, prec)
    inf = InferenceLDM(m, device=dev)
    pcs = _data.synthetic_clouds(n_obj, 1024, seed=1234, dist="S").to(dev)
    metas = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in default_metas(n_obj).items()}
    ms = timed(lambda: inf.generate_grasps(pcs, metas, num_grasps=grasps, seed=1), iters=3 if n_obj * grasps > 50000 else 5)
    flops = n_obj * 8.115e9 + n_obj * grasps * (steps * 7.589e6 + 30.70e6)
    return ms, n_obj * grasps / (ms * 1e-3), flops / (ms * 1e-3) / PEAK


print("# Other BASELINE.json configurations on one B200 (bf16 tensor-core path unless noted)\n")
print("`python tools/bench_configs.py`.  One batch at a time on one stream (no pipelining; the sampler picks one or two sample sets per CTA from the batch size), inputs resident in HBM, CUDA events, in-kernel Philox noise.  \"fraction of peak\" = algorithmic FLOPs (SURVEY.md 8d) / time / 1392.3 TFLOP/s (measured sustained bf16).\n")
print("| config | workload | ms / batch | grasps/s (clouds/s) | fraction of sustained bf16 peak |")
print("|---|---|---:|---:|---:|")
ms, r, f = ldm("fpc", "ddpm", 100, 1, 20)
print(f"| 1-like | LDM 100 DDPM steps, 1 object x 20 grasps (latency) | {ms:.2f} | {r:,.0f} | {f:.4f} |")
ms, r, f = ldm("fpc", "ddpm", 100, 64, 20)
print(f"| 2 | LDM 100 DDPM steps, 64 objects x 20 grasps | {ms:.2f} | {r:,.0f} | {f:.4f} |")
ms, r, f = ldm("fpc", "ddpm", 100, 64, 20, prec="fp32")
print(f"| 2 (fp32 SIMT path) | same, strict-fp32 kernels | {ms:.2f} | {r:,.0f} | {f:.4f} |")
for steps in (10, 50):
    ms, r, f = ldm("fpc", "ddim", steps, 128, 100)
    print(f"| 3 (per GPU) | LDM DDIM {steps} steps, 128 objects x 100 grasps (1024 objects over 8 GPUs) | {ms:.2f} | {r:,.0f} | {f:.4f} |")
for n_obj in (16, 64, 256):
    ms, r, f = ldm("fpc", "ddpm", 100, n_obj, 256)
    print(f"| 5 | LDM 100 DDPM steps, {n_obj} objects x 256 grasps | {ms:.2f} | {r:,.0f} | {f:.4f} |")
ms, r, f = ldm("ppc", "ddpm", 100, 64, 20)
print(f"| 4-like | ppc LDM (16-position latent) 100 DDPM steps, 64 objects x 20 grasps | {ms:.2f} | {r:,.0f} | - |")
# config 4: encoder only, partial-point-cloud config, 4096 clouds
m = _models.build("ppc").to(dev)
set_precision(m, "bf16")
pcs = _data.synthetic_clouds(4096, 1024, seed=1234, dist="S").to(dev)
ms = timed(lambda: m.vae_model.encode_pc(pcs), iters=3, warm=1)
print(f"| 4 | ppc encoder only, 4096 clouds x 1024 points | {ms:.2f} | {4096 / (ms * 1e-3):,.0f} clouds/s | {4096 * 8.115e9 / (ms * 1e-3) / PEAK:.4f} |")
