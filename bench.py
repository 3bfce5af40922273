#!/usr/bin/env python
"""Headline benchmark: grasps/sec of full LDM grasp generation (PVCNN encoder + T latent-denoising steps + grasp
decoder + pose post-processing) on synthetic point clouds - BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5]

--config selects a BASELINE.json configuration (default 2, the one the metric is quoted on):
  1  fpc VAE-mode generation, 1 cloud x 20 grasps per GPU
  2  fpc LDM, 100 DDPM steps, 64 objects x 20 grasps per GPU (weak scaling)
  3  fpc LDM, DDIM --ddim-steps 10|50, 1024 objects x 100 grasps in total, sharded by object (strong scaling)
  4  ppc encoder only, 4096 clouds per GPU (clouds/s)
  5  fpc LDM, 100 DDPM steps, --objects O x 256 grasps per GPU (weak scaling; default O = 256)

One "step" = one pass of the path over one batch.  `value` is timed with the point clouds resident in HBM; `e2e` goes
through graspldm_b200.inference.Inference{LDM,VAE}.generate_grasps (config 4: PVCNNEncoder.forward) with pinned HOST
buffers (H2D of the clouds and D2H of the results inside the timed region).  N > 1: one rank per GPU (torchrun), objects
sharded by rank, no collective on the compute path, one final all_gather of the packed results.  `--impl reference` times
the CPU oracle port (the reference's own python cannot travel to the GPU box and its encoder has no CPU path at all,
SURVEY.md finding 6) on rank 0, on a bounded sample of the same configuration.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("TQDM_DISABLE", "1")

import torch  # noqa: E402

N_POINTS = 1024
# algorithmic work (SURVEY.md 8d / BASELINE.md section 2), FLOPs
F_ENCODER_PER_CLOUD = 8.115e9
# the bf16 path composes conv_downscale (2.416 GFLOP) and out_layer.0 (5 MFLOP) into one [3, 1536] projection (9.4 MFLOP):
# FLOPs the kernels actually execute per cloud; the roofline keeps the reference's ALGORITHMIC count and states both
F_ENCODER_EXECUTED_PER_CLOUD = 8.115e9 - 2.416e9 - 5.1e6 + 2 * 3 * 1536 * 1024
F_DENOISER_PER_SAMPLE_STEP = 7.589e6
F_DECODER_PER_GRASP = 30.70e6


def workload(args, world):
    """The configuration both arms run: everything BASELINE.json fixes, nothing about how either arm executes it."""
    c = args.config
    if c == 1:
        w = dict(id=1, model="fpc", mode="vae", objects_per_gpu=1, grasps=20, T=0, sched=None, scaling="weak",
                 metric="grasps/sec (VAE mode)", unit="grasps/s",
                 text="fpc_1a_latentc3_z4_pc64 VAE-mode generation, 1 synthetic point cloud x 20 grasps per GPU (BASELINE.json configs[0])")
    elif c == 2:
        w = dict(id=2, model="fpc", mode="ldm", objects_per_gpu=64, grasps=20, T=100, sched="ddpm", scaling="weak",
                 metric="grasps/sec (LDM 100 steps)", unit="grasps/s",
                 text="fpc_1a_latentc3_z4_pc64 LDM-mode generation, 100 DDPM steps, 64 objects x 20 grasps per GPU "
                      "(BASELINE.json configs[1])")
    elif c == 3:
        w = dict(id=3, model="fpc", mode="ldm", objects_total=1024, grasps=100, T=args.ddim_steps, sched="ddim", scaling="strong",
                 metric=f"grasps/sec (LDM DDIM {args.ddim_steps} steps)", unit="grasps/s",
                 text=f"fpc LDM-mode DDIM {args.ddim_steps}-step sampling, 1024 objects x 100 grasps in total, sharded by object "
                      "(BASELINE.json configs[2])")
    elif c == 4:
        w = dict(id=4, model="ppc", mode="encoder", objects_per_gpu=4096, grasps=0, T=0, sched=None, scaling="weak",
                 metric="clouds/sec (ppc PVCNN encoder)", unit="clouds/s",
                 text="partial-point-cloud encoder config, 4096 clouds per GPU, encoder-only throughput (BASELINE.json configs[3])")
    else:
        w = dict(id=5, model="fpc", mode="ldm", objects_per_gpu=args.objects, grasps=256, T=100, sched="ddpm", scaling="weak",
                 metric="grasps/sec (LDM 100 steps)", unit="grasps/s",
                 text=f"fpc LDM-mode generation, 100 DDPM steps, {args.objects} objects x 256 grasps per GPU "
                      "(BASELINE.json configs[4], large-batch sweep)")
    w["objects"] = w["objects_total"] if "objects_total" in w else w["objects_per_gpu"] * world
    w["config"] = {"workload": w["text"] + ", random-init weights, 1024-point synthetic clouds", "config_id": w["id"],
                   "objects": w["objects"], "grasps_per_object": w["grasps"], "denoising_steps": w["T"], "scheduler": w["sched"],
                   "parallelism": f"objects sharded over {world} rank(s), one final all_gather"}
    return w


def units_per_step(w, n_obj):
    return n_obj if w["mode"] == "encoder" else n_obj * w["grasps"]


def flops(w, n_obj):
    """algorithmic FLOPs of one step over n_obj objects: (encoder, sampler, decoder)"""
    n = n_obj * w["grasps"]
    return (n_obj * F_ENCODER_PER_CLOUD, n * w["T"] * F_DENOISER_PER_SAMPLE_STEP if w["mode"] == "ldm" else 0.0,
            n * F_DECODER_PER_GRASP if w["mode"] != "encoder" else 0.0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"], tflops_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed regions (B200_PROFILING.md recipe).  Uses NVML in-process (a light
    query, every 20 ms over all timed passes) and falls back to polling nvidia-smi, which the recipe names."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        bits = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
        self.rows.append([str(sm), str(mx), f"{pw:.1f}"] + [("Active" if r & b else "Not Active") for _, b in bits])

    def run(self):
        while not self._halt.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                       capture_output=True, text=True, timeout=5).stdout.strip()
                    if o:
                        self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self._halt.wait(0.02 if self.nvml is not None else 0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        pw = [float(r[2]) for r in self.rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(max(mx) if mx else None), reasons=reasons,
                    samples=len(self.rows), power_w_max=(max(pw) if pw else None),
                    source="nvml" if self.nvml is not None else "nvidia-smi")


# --------------------------------------------------------------------------------------------------------------------
# CPU legs (the oracle port; test infrastructure used here only as the measured CPU baseline)
# --------------------------------------------------------------------------------------------------------------------
def cpu_sample_objects(w):
    """Objects of the bounded CPU sample: about 32 k sample-steps (16 objects x 20 grasps x 100 steps for config 2)."""
    if w["mode"] == "ldm":
        return max(1, min(64, 32000 // max(1, w["grasps"] * w["T"])))
    return 16


def make_cpu_step(w, n_obj):
    import _data
    import _models
    from oracle import model_torch as M
    model = _models.build(w["model"], scheduler=w["sched"] or "ddpm")
    vae_sd, ddm_sd = _models.split_state_dicts(model)
    pcs = _data.synthetic_clouds(n_obj, N_POINTS, seed=1234, dist="S")
    metas = dict(pc_mean=torch.zeros(n_obj, 3), pc_std=torch.full((n_obj, 3), 0.05), grasp_mean=torch.zeros(1, 6),
                 grasp_std=torch.tensor([[.05, .05, .05, .5, .5, .5]]))
    D = 4 if w["model"] == "fpc" else 16
    G = w["grasps"]

    def one():
        g = torch.Generator().manual_seed(42)
        with torch.no_grad():
            if w["mode"] == "encoder":
                return M.pvcnn_encoder_forward(vae_sd, "encoder.pc_encoder.", pcs)
            if w["mode"] == "vae":
                tm, lg = M.generate_grasps_vae(vae_sd, pcs, G, torch.randn(n_obj * G, D, generator=g))
            else:
                x_T = torch.randn(n_obj * G, 1, D, generator=g)
                tm, lg = M.generate_grasps_ldm(vae_sd, ddm_sd, pcs, G, x_T, num_inference_steps=w["T"], kind=w["sched"])
            return M.postprocess(tm, lg, pcs, metas, n_obj, G)
    return one


def cpu_baseline_sample(w):
    """Bounded CPU sample of the same workload on the host cores (reported beside the GPU number): warm-up 1, median of
    3 passes with every host thread, one more pass with 8 threads."""
    cores = os.cpu_count() or 1
    n_obj = cpu_sample_objects(w)
    one = make_cpu_step(w, n_obj)
    units = units_per_step(w, n_obj)
    torch.set_num_threads(cores)
    one()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        one()
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    torch.set_num_threads(min(8, cores))
    t0 = time.perf_counter()
    one()
    t8 = time.perf_counter() - t0
    torch.set_num_threads(cores)
    return {"value": units / med, "unit": w["unit"], "cores": cores, "kind": "port",
            "sample": f"{n_obj} objects x {w['grasps']} grasps, {w['T']} steps per pass; warm-up 1, median of 3 passes "
                      f"({med:.2f} s each) on {cores} threads",
            "value_8_threads": units / t8, "passes_s": [round(t, 3) for t in ts]}


def run_reference(args, rank, world):
    """CPU arm: the oracle port on the host cores, all threads, a bounded sample of the arm's configuration per step."""
    if rank != 0:
        return
    w = workload(args, world)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_obj = cpu_sample_objects(w)
    one = make_cpu_step(w, n_obj)
    t0 = time.perf_counter()
    one()                                                   # calibration pass (also the first warm-up step)
    t1 = time.perf_counter() - t0
    budget = 240.0                                          # the whole run stays within a few minutes
    if t1 * (args.steps + args.warmup) > budget and n_obj > 2:
        n_obj = max(2, int(n_obj * budget / (t1 * (args.steps + args.warmup))))
        one = make_cpu_step(w, n_obj)
        one()
    for _ in range(max(0, args.warmup - 1)):
        one()
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        one()
        ts.append(time.perf_counter() - t0)
    dt = sum(ts) / len(ts)
    units = units_per_step(w, n_obj)
    v = units / dt
    sample = (f"{n_obj} objects x {w['grasps']} grasps, {w['T']} steps per step ({args.steps} steps timed after {args.warmup} "
              f"warm-up, median step {statistics.median(ts):.2f} s)")
    print(json.dumps({
        "impl": "reference", "metric": w["metric"], "value": v, "unit": w["unit"], "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": w["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": w["config"],
        "cpu_baseline": {"value": v, "unit": w["unit"], "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": w["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


# --------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)     # config 2: 200 x ~5 ms = a 1 s timed region
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json configuration (see the module docstring)")
    ap.add_argument("--ddim-steps", type=int, default=10, choices=[10, 50], help="config 3: DDIM steps")
    ap.add_argument("--objects", type=int, default=256, help="config 5: objects per GPU (x 256 grasps)")
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"],
                    help="bf16: tcgen05 tensor-core kernels (bf16 operands, fp32 accumulation); fp32: strict-fp32 SIMT parity path")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=5,
                    help="batches in flight: consecutive steps are issued round-robin on this many CUDA streams, so the "
                         "encoder / decoder of one batch fill the SMs the persistent sampler of another leaves idle")
    args = ap.parse_args()
    # the contract is ONE JSON line on stdout: whatever native libraries print there (NCCL's version banner ...) goes to
    # stderr instead; file descriptor 1 is restored for the final line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(obj), flush=True)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    import _data
    import _models
    from graspldm_b200 import _lib, engine, sharding
    from graspldm_b200.inference import InferenceLDM, InferenceVAE, default_metas

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    w = workload(args, world)
    if os.environ.get("GLDM_TC_SETS"):
        _lib.call("gldm_sampler_tc_set_sets", int(os.environ["GLDM_TC_SETS"]))
    model = _models.build(w["model"], scheduler=w["sched"] or "ddpm").to(dev)
    if w["mode"] == "ldm":
        model.set_inference_timesteps(w["T"])
    model.diffusion_model.rng_mode = "fused"          # noise drawn inside the sampler kernel (Philox4x32-10)
    model.diffusion_model.precision = args.precision
    model.vae_model.encoder.pc_encoder.precision = args.precision
    model.vae_model.decoder.precision = args.precision
    inf = InferenceLDM(model, device=dev) if w["mode"] == "ldm" else InferenceVAE(model.vae_model, device=dev)
    n_total, G = w["objects"], w["grasps"]
    lo, hi = sharding.shard_bounds(n_total, world, rank)
    n_loc = hi - lo
    counts = sharding.shard_counts(n_total, world)
    # clouds: distinct up to 256 objects per rank, then repeated (the kernels do not care; host memory does)
    base = _data.synthetic_clouds(min(n_loc, 256), N_POINTS, seed=1234 + rank, dist="S")
    pcs_host = base.repeat((n_loc + base.shape[0] - 1) // base.shape[0], 1, 1)[:n_loc].contiguous().pin_memory()
    metas = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in default_metas(n_loc).items()}
    metas_dev = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in metas.items()}
    pcs_dev = pcs_host.to(dev)
    n_streams = max(1, args.streams)
    chunk = 1024 if w["mode"] == "encoder" else max(1, min(n_loc, 131072 // max(1, G)))     # objects per generation call
    if w["mode"] == "encoder":
        F_out = 64 if w["model"] == "fpc" else 256
        out_hosts = [{"z_pc": torch.empty((n_loc, 3, F_out)).pin_memory()} for _ in range(n_streams)]
    else:
        out_hosts = [{"grasps": torch.empty((n_loc, G, 4, 4)).pin_memory(),
                      "confidence": torch.empty((n_loc, G, 1)).pin_memory()} for _ in range(n_streams)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    enc = model.vae_model.encoder.pc_encoder

    def generate(pcs, mt, seed):
        """the public call(s) of one step over this rank's objects -> dict of per-object results"""
        outs = []
        for s in range(0, n_loc, chunk):
            e = min(n_loc, s + chunk)
            if w["mode"] == "encoder":
                outs.append({"z_pc": enc(pcs[s:e].to(dev, non_blocking=True))})
                continue
            m = {k: (v[s:e] if isinstance(v, torch.Tensor) and v.shape[0] == n_loc else v) for k, v in mt.items()}
            kw = dict(seed=seed) if w["mode"] == "ldm" else {}
            o = inf.generate_grasps(pcs[s:e], m, num_grasps=G, **kw)
            outs.append({"grasps": o["grasps"], "confidence": o["confidence"]})
        return outs[0] if len(outs) == 1 else {k: torch.cat([o[k] for o in outs]) for k in outs[0]}

    def gen_resident(seed):
        out = generate(pcs_dev, metas_dev, seed)
        if world > 1:
            out = sharding.gather_results(out, counts)
        return out

    def gen_e2e(seed):
        oh = out_hosts[seed % n_streams]
        out = generate(pcs_host, metas, seed)                 # H2D inside
        for k, v in out.items():
            oh[k].copy_(v, non_blocking=True)                 # D2H of the result
        if world > 1:
            sharding.gather_results(out, counts)
        return out

    def fence():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def timed(fn, steps, warmup):
        """sequential pass, one batch at a time: per-batch latency (L2 flushed between iterations, outside the events)"""
        import gc
        for i in range(warmup):
            fn(i)
        gc.collect()
        gc.disable()      # a generational collection inside the timed loop stalls the launching thread for tens of ms
        fence()
        evs = []
        engine.SECTIONS.enabled = True
        engine.SECTIONS.events = []
        for i in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(1000 + i)
            b.record()
            evs.append((a, b))
        fence()
        engine.SECTIONS.enabled = False
        gc.enable()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in evs)), engine.SECTIONS.collect()

    def timed_pipelined(fn, steps, warmup, sections):
        """K steps issued round-robin on n_streams streams; one event pair brackets the whole region.  The L2 flush
        of every step is inside the timed region here (it runs on the step's own stream).  With `sections` every
        encoder / sampler / decoder launch is also bracketed by events on its own stream."""
        import gc
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        main = torch.cuda.current_stream(dev)
        for i in range(max(warmup, n_streams)):
            with torch.cuda.stream(streams[i % n_streams]):
                fn(i)
        gc.collect()
        gc.disable()
        fence()
        engine.SECTIONS.enabled = sections
        engine.SECTIONS.events = []
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main)
        for st in streams:
            st.wait_event(a)
        for i in range(steps):
            with torch.cuda.stream(streams[i % n_streams]):
                flush.zero_()
                fn(1000 + i)
        for st in streams:
            main.wait_stream(st)
        b.record(main)
        fence()
        engine.SECTIONS.enabled = False
        gc.enable()
        return max_over_ranks(a.elapsed_time(b)), engine.SECTIONS.collect()

    tc = args.precision == "bf16"
    rows_env = os.environ.get("GLDM_TC_ROWS")
    clocks = ClockSampler(local_rank)
    clocks.start()
    # pass 1 (sequential, one batch at a time): per-batch latency with the automatic kernel choice
    lat_steps = max(3, min(args.steps, 20))
    lat_ms, lat_sections = timed(gen_resident, lat_steps, args.warmup)
    # pass 2 (the reported value): batches in flight on n_streams streams.  Throughput mode of the sampler: the row-major
    # tcgen05 kernel (32 samples per CTA) - forced here, the automatic choice takes it only above one wave of CTAs
    l0 = _lib.launch_count()
    if n_streams > 1:
        if tc and rows_env is None:
            _lib.call("gldm_sampler_tc_set_rows", 1)
        total_ms, sections = timed_pipelined(gen_resident, args.steps, args.warmup, sections=True)
        launches = (_lib.launch_count() - l0) // (args.steps + max(args.warmup, n_streams)) * args.steps
        e2e_ms, _ = timed_pipelined(gen_e2e, args.steps, args.warmup, sections=False)
    else:
        total_ms, sections = timed(gen_resident, args.steps, args.warmup)
        launches = (_lib.launch_count() - l0) // (args.steps + args.warmup) * args.steps
        e2e_ms, _ = timed(gen_e2e, args.steps, args.warmup)
    clk = clocks.finish()

    ms_per_step = total_ms / args.steps
    units = units_per_step(w, n_total)
    value = units / (ms_per_step * 1e-3)
    e2e_value = units / (e2e_ms / args.steps * 1e-3)

    pk = peaks()
    mean = lambda xs: (sum(xs) / len(xs)) if xs else 0.0
    f_enc, f_samp, f_dec = flops(w, n_loc)
    calls = max(1, -(-n_loc // chunk))                    # generation calls (= launches of each section) per step
    rows_kernel = tc and w["mode"] == "ldm" and w["model"] == "fpc" and (n_streams > 1 and rows_env is None or rows_env == "1"
                                                                         or chunk * G > 16 * 148)
    names = {"encoder": ("PVCNN encoder pass: conv3d_tc3p / tc3 / tc16 kernels, gemm_tc_kernel (tcgen05) + SIMT voxel glue"
                         if tc else "PVCNN encoder pass (fp32 SIMT)"),
             "sampler": (("rows::resnet_rows_kernel" if rows_kernel else "resnet_tc_kernel<L,NSETS>") +
                         " (tcgen05 persistent T-step sampler, one launch per call)") if tc else "resnet_kernel<L> (fp32 SIMT persistent sampler)",
             "decoder": "rows::resnet_rows_kernel<16> (tcgen05 grasp decoder, row-major, 8 grasps per CTA)" if tc else "resnet_kernel<16> (fp32 SIMT decoder)"}
    kernels = []
    for sec, fl in (("encoder", f_enc), ("sampler", f_samp), ("decoder", f_dec)):
        ms = mean(sections.get(sec, []))
        if fl > 0 and ms > 0:
            tf = fl / calls / (ms * 1e-3) / 1e12
            kernels.append({"section": sec, "name": names[sec], "ms_in_timed_region": ms, "launches_per_step": calls,
                            "algorithmic_flops_per_launch": fl / calls, "achieved_tflops": tf, "frac": tf / pk["tflops_sustained"],
                            "ms_alone": mean(lat_sections.get(sec, []))})
    folded = tc and os.environ.get("GLDM_FOLD_DOWNSCALE", "1") != "0"
    f_enc_exec = f_enc * (F_ENCODER_EXECUTED_PER_CLOUD / F_ENCODER_PER_CLOUD) if folded else f_enc
    for k in kernels:
        if k["section"] == "encoder" and folded:
            k["executed_flops_per_launch"] = f_enc_exec / calls
            k["executed_tflops"] = f_enc_exec / calls / (k["ms_in_timed_region"] * 1e-3) / 1e12
            k["note"] = ("conv_downscale + out_layer.0 are composed into one projection applied in the epilogue of the 768->1536 "
                         "GEMM: 5.70 of the reference's 8.115 GFLOP per cloud are executed; frac uses the algorithmic count")
    dom = max(kernels, key=lambda k: k["ms_in_timed_region"] * k["launches_per_step"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r02_tc_sampler_traffic.json")
    if dom["section"] == "sampler" and tc and w["id"] == 2 and os.path.exists(tp):      # the capture is config 2's launch
        t = json.load(open(tp))["resnet_rows_kernel"]          # dram__bytes_read + write of one ncu --set full capture
        traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    # the persistent sampler holds ONE CTA per SM for its whole launch and fills only ceil(samples / samples-per-CTA) SMs:
    # its rate against the peak of the SMs it actually holds (the rest of the GPU runs the other batches' kernels meanwhile)
    for k in kernels:
        if k["section"] == "sampler" and tc and w["model"] == "fpc":
            per_cta = 32 if rows_kernel else 16
            ctas = min(148, -(-(min(chunk, n_loc) * G) // per_cta))
            k["ctas"] = ctas
            k["frac_of_held_sms"] = k["achieved_tflops"] / (pk["tflops_sustained"] * ctas / 148.0)
    total_flops = f_enc + f_samp + f_dec
    roofline = {"bound": "tensor", "kernel": dom["name"], "achieved": dom["achieved_tflops"], "peak": pk["tflops_sustained"],
                "unit": "TFLOP/s", "frac": dom["frac"], "traffic": traffic, "peak_source": pk["source"] + " sustained bf16",
                "algorithmic_flops_per_launch": dom["algorithmic_flops_per_launch"], "kernel_ms": dom["ms_in_timed_region"],
                "how": "CUDA events on the launching stream around every launch of the section INSIDE the timed region; with "
                       f"{n_streams} batches in flight the kernel shares the GPU with the other batches' kernels, so its duration "
                       "there is longer than alone (ms_alone: the sequential latency pass)",
                "kernels": kernels,
                # all algorithmic FLOPs of a step over the whole timed region: what the GPU sustains end to end
                "whole_step": {"algorithmic_flops_per_step": total_flops,
                               "achieved_tflops": total_flops / (ms_per_step * 1e-3) / 1e12,
                               "frac": total_flops / (ms_per_step * 1e-3) / 1e12 / pk["tflops_sustained"],
                               "executed_flops_per_step": f_enc_exec + f_samp + f_dec,
                               "executed_frac": (f_enc_exec + f_samp + f_dec) / (ms_per_step * 1e-3) / 1e12 / pk["tflops_sustained"]}}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None if (args.no_cpu_baseline or world > 1) else cpu_baseline_sample(w)
    h2d = pcs_host.numel() * 4 + (n_loc * G * (4 if w["model"] == "fpc" else 16) * 4 if w["mode"] != "encoder" else 0)
    line = {
        "metric": w["metric"], "value": value, "unit": w["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
        "dtype": "bf16" if tc else "f32", "data": "synthetic", "config": w["config"],
        "run": {"l2": "256 MiB buffer written before every step" + (" (inside the timed region, on the step's stream)" if n_streams > 1 else " (between timed iterations)"),
                "batches_in_flight": n_streams, "objects_per_call": chunk, "precision": args.precision,
                "latency_ms_per_batch": lat_ms / lat_steps,
                "latency_pass": "one batch at a time, automatic sampler kernel choice (channel-major 16-sample CTAs up to 2368 samples)",
                "sampler_kernel_timed_region": names["sampler"] if w["mode"] == "ldm" else None,
                "rng": "in-kernel Philox4x32-10 + Box-Muller (x_T drawn on the host generator as the reference does)"},
        "e2e": {"value": e2e_value, "unit": w["unit"], "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(sum(v.numel() * 4 for v in out_hosts[0].values())), "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
