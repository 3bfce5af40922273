#!/usr/bin/env python
"""Headline benchmark: grasps/sec of full LDM grasp generation (PVCNN encoder + 100 DDPM latent-denoising
steps + grasp decoder + pose post-processing) on synthetic point clouds - BASELINE.json's metric on its
config 2 (fpc_1a_latentc3_z4_pc64, 64 objects x 20 grasps per GPU, random-init weights).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one generation pass over one batch.  `value` is timed with the point clouds resident in HBM;
`e2e` goes through graspldm_b200.inference.InferenceLDM.generate_grasps with pinned HOST buffers (H2D of
the clouds and D2H of poses + confidences inside the timed region).  N > 1: one rank per GPU (torchrun),
objects sharded by rank, no collective on the compute path, one final all_gather of the results
(weak scaling: 64 objects per GPU).  `--impl reference` times the CPU oracle port (the reference's own
python cannot travel to the GPU box and its encoder has no CPU path at all, SURVEY.md finding 6) on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("TQDM_DISABLE", "1")

import torch  # noqa: E402

N_OBJ, N_GRASPS, N_STEPS_DDPM, N_POINTS = 64, 20, 100, 1024
METRIC, UNIT = "grasps/sec (LDM 100 steps)", "grasps/s"
# algorithmic work (SURVEY.md 8d / BASELINE.md section 2), FLOPs
F_ENCODER_PER_CLOUD = 8.115e9
F_DENOISER_PER_SAMPLE_STEP = 7.589e6
F_DECODER_PER_GRASP = 30.70e6


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"], tflops_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md recipe).  Uses NVML in-process
    (a light query) and falls back to polling nvidia-smi, which the recipe names."""
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
            self._halt.wait(0.1 if self.nvml is not None else 0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(max(mx) if mx else None), reasons=reasons,
                    samples=len(self.rows), source="nvml" if self.nvml is not None else "nvidia-smi")


def run_reference(args, rank, world):
    """CPU arm: the oracle port on the host cores, all threads, bounded sample per step."""
    if rank != 0:
        return
    import _data
    import _models
    from oracle import model_torch as M
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_obj = 2
    model = _models.build("fpc")
    vae_sd, ddm_sd = _models.split_state_dicts(model)
    pcs = _data.synthetic_clouds(n_obj, seed=1234, dist="S")
    metas = dict(pc_mean=torch.zeros(n_obj, 3), pc_std=torch.full((n_obj, 3), 0.05), grasp_mean=torch.zeros(1, 6),
                 grasp_std=torch.tensor([[.05, .05, .05, .5, .5, .5]]))

    def one():
        g = torch.Generator().manual_seed(42)
        x_T = torch.randn(n_obj * N_GRASPS, 1, 4, generator=g)
        with torch.no_grad():
            tm, lg = M.generate_grasps_ldm(vae_sd, ddm_sd, pcs, N_GRASPS, x_T, num_inference_steps=N_STEPS_DDPM)
            return M.postprocess(tm, lg, pcs, metas, n_obj, N_GRASPS)

    for _ in range(min(args.warmup, 1)):
        one()
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = (time.perf_counter() - t0) / steps
    v = n_obj * N_GRASPS / dt
    sample = f"{n_obj} objects x {N_GRASPS} grasps, {N_STEPS_DDPM} DDPM steps per step ({steps} steps timed)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "fpc_1a_latentc3_z4_pc64 LDM 100 DDPM steps, CPU oracle port (plain PyTorch fp32 + numpy ops)",
                   "objects_per_step": n_obj, "grasps_per_object": N_GRASPS},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def cpu_baseline_sample():
    """Bounded CPU sample of the same workload on the host cores (reported beside the GPU number)."""
    import _data
    import _models
    from oracle import model_torch as M
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_obj = 2
    model = _models.build("fpc")
    vae_sd, ddm_sd = _models.split_state_dicts(model)
    pcs = _data.synthetic_clouds(n_obj, seed=1234, dist="S")
    g = torch.Generator().manual_seed(42)
    x_T = torch.randn(n_obj * N_GRASPS, 1, 4, generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        M.generate_grasps_ldm(vae_sd, ddm_sd, pcs, N_GRASPS, x_T, num_inference_steps=N_STEPS_DDPM)
    dt = time.perf_counter() - t0
    return {"value": n_obj * N_GRASPS / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_obj} objects x {N_GRASPS} grasps, {N_STEPS_DDPM} DDPM steps, 1 pass, {dt:.1f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)     # 50 x ~6 ms: the four-deep pipeline's fill / drain is < 3 % of it
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"],
                    help="bf16: tcgen05 tensor-core kernels (bf16 operands, fp32 accumulation); fp32: strict-fp32 SIMT parity path")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=4,
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
    from graspldm_b200.inference import InferenceLDM, default_metas

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    if os.environ.get("GLDM_TC_SETS"):
        _lib.call("gldm_sampler_tc_set_sets", int(os.environ["GLDM_TC_SETS"]))
    model = _models.build("fpc").to(dev)
    model.set_inference_timesteps(N_STEPS_DDPM)
    model.diffusion_model.rng_mode = "fused"          # noise drawn inside the sampler kernel (Philox4x32-10)
    model.diffusion_model.precision = args.precision
    model.vae_model.encoder.pc_encoder.precision = args.precision
    model.vae_model.decoder.precision = args.precision
    inf = InferenceLDM(model, device=dev)
    n_total = N_OBJ * world                            # weak scaling: 64 objects per GPU
    lo, hi = sharding.shard_bounds(n_total, world, rank)
    pcs_host = _data.synthetic_clouds(hi - lo, N_POINTS, seed=1234 + rank, dist="S").pin_memory()
    metas = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in default_metas(hi - lo).items()}
    metas_dev = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in metas.items()}
    pcs_dev = pcs_host.to(dev)
    counts = sharding.shard_counts(n_total, world)
    n_streams = max(1, args.streams)
    out_hosts = [{"grasps": torch.empty((hi - lo, N_GRASPS, 4, 4)).pin_memory(),
                  "confidence": torch.empty((hi - lo, N_GRASPS, 1)).pin_memory()} for _ in range(n_streams)]
    out_host = out_hosts[0]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def gen_resident(seed):
        out = inf.generate_grasps(pcs_dev, metas_dev, num_grasps=N_GRASPS, seed=seed)
        if world > 1:
            out = sharding.gather_results({"grasps": out["grasps"], "confidence": out["confidence"]}, counts)
        return out

    def gen_e2e(seed):
        oh = out_hosts[seed % n_streams]
        out = inf.generate_grasps(pcs_host, metas, num_grasps=N_GRASPS, seed=seed)   # H2D inside
        oh["grasps"].copy_(out["grasps"], non_blocking=True)                           # D2H of the result
        oh["confidence"].copy_(out["confidence"], non_blocking=True)
        if world > 1:
            sharding.gather_results({"grasps": out["grasps"], "confidence": out["confidence"]}, counts)
        return out

    def timed(fn, steps, warmup, sections=False):
        import gc
        for i in range(warmup):
            fn(i)
        gc.collect()
        gc.disable()      # a generational collection inside the timed loop stalls the launching thread for tens of ms
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        evs = []
        engine.SECTIONS.enabled = sections
        engine.SECTIONS.events = []
        for i in range(steps):
            flush.zero_()                                  # L2 flush between timed iterations (outside the events)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(1000 + i)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        engine.SECTIONS.enabled = False
        gc.enable()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, engine.SECTIONS.collect()

    def timed_pipelined(fn, steps, warmup):
        """K steps issued round-robin on n_streams streams; one event pair brackets the whole region.  The L2 flush
        of every step is inside the timed region here (it runs on the step's own stream)."""
        import gc
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        main = torch.cuda.current_stream(dev)
        for i in range(max(warmup, n_streams)):
            with torch.cuda.stream(streams[i % n_streams]):
                fn(i)
        gc.collect()
        gc.disable()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
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
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        gc.enable()
        total_ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms

    # pass 1 (sequential, one batch at a time): per-batch latency and the per-section / per-kernel durations
    lat_ms, sections = timed(gen_resident, args.steps, args.warmup, sections=True)
    # the sampler kernel of the pipelined passes (row-major, 32 samples per CTA), timed alone the same way
    rows_ms = None
    if n_streams > 1 and args.precision == "bf16" and os.environ.get("GLDM_TC_ROWS") is None:
        _lib.call("gldm_sampler_tc_set_rows", 1)
        _, sec_rows = timed(gen_resident, max(3, args.steps // 2), 2, sections=True)
        _lib.call("gldm_sampler_tc_set_rows", -1)
        rows_ms = sum(sec_rows.get("sampler", [0.0])) / max(1, len(sec_rows.get("sampler", [])))
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = _lib.launch_count()
    if n_streams > 1:
        # throughput mode: 32 samples per sampler CTA in two interleaved sets (fewer SMs per batch, the UMMA phase of
        # one set under the epilogue of the other); the sequential latency pass above used the automatic choice
        if args.precision == "bf16" and not os.environ.get("GLDM_TC_SETS"):
            _lib.call("gldm_sampler_tc_set_sets", 2)
        if args.precision == "bf16" and os.environ.get("GLDM_TC_ROWS") is None:
            _lib.call("gldm_sampler_tc_set_rows", 1)     # throughput mode: the row-major sampler kernel (32 samples per CTA)
        total_ms = timed_pipelined(gen_resident, args.steps, args.warmup)
        launches = (_lib.launch_count() - l0) // (args.steps + max(args.warmup, n_streams)) * args.steps
    else:
        total_ms, _ = timed(gen_resident, args.steps, args.warmup)
        launches = (_lib.launch_count() - l0) // (args.steps + args.warmup) * args.steps
    clk = clocks.finish()
    e2e_ms = timed_pipelined(gen_e2e, args.steps, args.warmup) if n_streams > 1 else timed(gen_e2e, args.steps, args.warmup)[0]

    ms_per_step = total_ms / args.steps
    grasps_per_step = n_total * N_GRASPS
    value = grasps_per_step / (ms_per_step * 1e-3)
    e2e_value = grasps_per_step / (e2e_ms / args.steps * 1e-3)

    pk = peaks()
    n_local = (hi - lo) * N_GRASPS
    samp_ms = sum(sections.get("sampler", [0.0])) / max(1, len(sections.get("sampler", [])))
    enc_ms = sum(sections.get("encoder", [0.0])) / max(1, len(sections.get("encoder", [])))
    dec_ms = sum(sections.get("decoder", [0.0])) / max(1, len(sections.get("decoder", [])))
    samp_flops = n_local * N_STEPS_DDPM * F_DENOISER_PER_SAMPLE_STEP
    achieved = samp_flops / (samp_ms * 1e-3) / 1e12 if samp_ms > 0 else 0.0
    kname = ("resnet_tc_kernel<4,1> in the latency pass / resnet_rows_kernel in the pipelined passes (tcgen05 persistent 100-step sampler, one launch per batch)" if args.precision == "bf16"
             else "resnet_kernel<4> (fp32 SIMT persistent 100-step sampler, one launch per batch)")
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_tc_sampler_traffic.json")
    if args.precision == "bf16" and os.path.exists(tp):
        t = json.load(open(tp))["sampler_tc_kernel"]          # dram__bytes_read + write of one ncu --set full capture
        traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    roofline = {"bound": "tensor", "kernel": kname,
                "achieved": achieved, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["tflops_sustained"], "traffic": traffic, "peak_source": pk["source"] + " sustained bf16",
                "algorithmic_flops_per_launch": samp_flops, "kernel_ms": samp_ms,
                # the same FLOPs over the whole pipelined timed region (which also holds the encoder / decoder kernels
                # of the batches in flight): a lower bound on what the sampler kernel sustains across the GPU
                "achieved_timed_region": samp_flops * args.steps / (total_ms * 1e-3) / 1e12,
                # the row-major kernel of the pipelined passes alone: one batch = ceil(samples / 32) CTAs, one per SM
                "pipelined_kernel": (None if not rows_ms else {
                    "name": "resnet_rows_kernel", "kernel_ms": rows_ms, "ctas": -(-n_local // 32), "sms": 148,
                    "achieved": samp_flops / (rows_ms * 1e-3) / 1e12,
                    "frac_of_peak_of_occupied_sms": samp_flops / (rows_ms * 1e-3) / 1e12 / (pk["tflops_sustained"] * min(1.0, -(-n_local // 32) / 148))}),
                "sections_ms": {"encoder": enc_ms, "sampler": samp_ms, "decoder": dec_ms},
                "note": ("sampler, decoder, encoder point-wise layers and Conv3d on tcgen05 (bf16 operands, fp32 accumulate); voxelize / devoxelize / GroupNorm+Swish / SE on fp32 SIMT kernels over channels-last grids; kernel_ms and sections_ms come from the sequential latency pass (channel-major sampler kernel, 16 samples per CTA, 80 CTAs; the pipelined passes use the row-major kernel, 32 samples per CTA)"
                         if args.precision == "bf16" else "strict-fp32 SIMT (FFMA) parity path")}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None if args.no_cpu_baseline else cpu_baseline_sample()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "fpc_1a_latentc3_z4_pc64 LDM-mode generation, 100 DDPM steps, 64 objects x 20 grasps per GPU "
                               "(BASELINE.json configs[1]), random-init weights, 1024-point synthetic clouds",
                   "objects": n_total, "grasps_per_object": N_GRASPS, "denoising_steps": N_STEPS_DDPM,
                   "parallelism": f"objects sharded over {world} rank(s), one final all_gather",
                   "l2": "256 MiB buffer written before every step" + (" (inside the timed region, on the step's stream)" if n_streams > 1 else " (between timed iterations)"),
                   "batches_in_flight": n_streams, "latency_ms_per_batch": lat_ms / args.steps, "precision": args.precision,
                   "sampler_samples_per_cta": ("32 (row-major tcgen05 kernel: rows = 4 positions x 32 samples) in the pipelined passes, 16 (channel-major kernel, 80 CTAs) in the latency pass"
                                               if (n_streams > 1 and args.precision == "bf16") else "automatic"),
                   "rng": "in-kernel Philox4x32-10 + Box-Muller (x_T drawn on the host generator as the reference does)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pcs_host.numel() * 4 + n_local * 4 * 4),
                "d2h_bytes_per_step": int(out_host["grasps"].numel() * 4 + out_host["confidence"].numel() * 4),
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
