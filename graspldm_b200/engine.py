"""Host-side orchestration of the CUDA generation path: weight packing (one-time) and kernel launches.

Everything numerical happens in libgraspldm_b200.so (include/graspldm_b200.h).  PyTorch is used for
device memory, streams and one-time layout work on weights (permutes, BatchNorm folding constants).
There is no CPU path: tensors must live on a CUDA device.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import GldmResNetCfg, GldmSamplerArgs

PRECISIONS = ("fp32", "bf16")   # "bf16": tcgen05 tensor-core kernels (bf16 operands, fp32 accumulation)


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


class _Sections:
    """Optional CUDA-event timing of the path's sections on the launching stream (bench.py turns it on to get
    the per-kernel durations the roofline is computed from; off by default, zero overhead)."""

    def __init__(self):
        self.enabled = False
        self.events = []

    def start(self, name, dev):
        if not self.enabled:
            return None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(torch.cuda.current_stream(dev))
        return (name, a, b, dev)

    def stop(self, tok):
        if tok is not None:
            tok[2].record(torch.cuda.current_stream(tok[3]))
            self.events.append(tok[:3])

    def collect(self):
        """-> {name: [ms, ...]} (call after a synchronize)"""
        out = {}
        for name, a, b in self.events:
            out.setdefault(name, []).append(a.elapsed_time(b))
        self.events = []
        return out


SECTIONS = _Sections()


def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (graspldm_b200 has no CPU fallback)")


# --------------------------------------------------------------------------------------------------
# ResNet1D family
# --------------------------------------------------------------------------------------------------
def make_resnet_cfg(L, dims, emb_dim, cond_ch, cond_dim, groups, time_cond, fourier_half):
    cfg = GldmResNetCfg()
    cfg.L = int(L)
    cfg.n_stages = len(dims) - 1
    for i, d in enumerate(dims):
        cfg.ch[i] = int(d)
    cfg.emb_dim, cfg.cond_ch, cfg.cond_dim = int(emb_dim), int(cond_ch), int(cond_dim)
    cfg.groups, cfg.time_cond, cfg.fourier_half = int(groups), int(bool(time_cond)), int(fourier_half)
    cfg.heads, cfg.dim_head = 4, 32
    return cfg


def resnet_param_keys(n_stages, time_cond):
    """Canonical order of the raw parameter blob (must match make_layout in csrc/resnet1d_f32.cu)."""
    def rb(p):
        return [p + "mlp.1.weight", p + "mlp.1.bias", p + "block1.proj.weight", p + "block1.proj.bias",
                p + "block1.norm.weight", p + "block1.norm.bias", p + "block2.proj.weight", p + "block2.proj.bias",
                p + "block2.norm.weight", p + "block2.norm.bias"]
    keys = ["init_conv.weight", "init_conv.bias"]
    if time_cond:
        keys += ["time_mlp.0.weights", "time_mlp.1.weight", "time_mlp.1.bias", "time_mlp.3.weight", "time_mlp.3.bias"]
    keys += ["input_emb_layers.0.weight", "input_emb_layers.0.bias"]
    for i in range(n_stages):
        b = f"blocks.{i}."
        keys += rb(b + "0.") + rb(b + "1.")
        keys += [b + "2.fn.norm.g", b + "2.fn.fn.to_qkv.weight", b + "2.fn.fn.to_out.0.weight",
                 b + "2.fn.fn.to_out.0.bias", b + "2.fn.fn.to_out.1.g", b + "3.weight", b + "3.bias"]
    keys += rb("final_res_block.") + ["final_conv.weight", "final_conv.bias"]
    return keys


def _flatten_padded(tensors, device):
    parts = []
    for t in tensors:
        f = t.detach().to(device=device, dtype=torch.float32).reshape(-1)
        pad = (-f.numel()) % 4
        parts.append(f)
        if pad:
            parts.append(torch.zeros(pad, device=device, dtype=torch.float32))
    return torch.cat(parts).contiguous()


def _signature(module):
    ps = list(module.parameters())
    return (ps[0].device, ps[0].data_ptr(), sum(p._version for p in ps))


class PackedResNet:
    def __init__(self, module, L, cond_ch):
        ps = list(module.parameters())
        dev = ps[0].device
        _require_cuda(ps[0], "model parameters")
        self.cfg = module.kernel_cfg(L, cond_ch)
        sd = module.state_dict()
        keys = resnet_param_keys(self.cfg.n_stages, self.cfg.time_cond)
        with torch.cuda.device(dev):
            raw = _flatten_padded([sd[k] for k in keys], dev)
            want = _lib.lib().gldm_resnet_raw_floats(ctypes.byref(self.cfg))
            if want != raw.numel():
                raise RuntimeError(f"resnet blob size mismatch: packed {raw.numel()} floats, library expects {want} "
                                   f"({_lib.lib().gldm_last_error().decode()})")
            self.prepared = torch.empty(_lib.lib().gldm_resnet_prepared_floats(ctypes.byref(self.cfg)),
                                        device=dev, dtype=torch.float32)
            _lib.call("gldm_resnet_prepare", ctypes.byref(self.cfg), raw.data_ptr(), self.prepared.data_ptr(),
                      _stream(dev))
        self.raw = raw
        self.device = dev
        self.L = L
        self._tc_pack = None
        self._tc_tables = {}

    def tc_tables(self, timesteps, coef):
        """Device-resident scheduler coefficients and time-embedding table of a schedule (cached per schedule)."""
        key = (tuple(int(t) for t in timesteps), coef.data_ptr(), coef._version)
        ent = self._tc_tables.get(key)
        if ent is None:
            dev = self.device
            with torch.cuda.device(dev):
                ts = torch.tensor(key[0], dtype=torch.int32, device=dev)
                cf = coef.detach().to(device=dev, dtype=torch.float32).contiguous()
                te = torch.empty((len(key[0]), self.cfg.emb_dim), device=dev, dtype=torch.float32)
                _lib.call("gldm_time_embed_table", ctypes.byref(self.cfg), self.raw.data_ptr(), ts.data_ptr(), len(key[0]),
                          te.data_ptr(), _stream(dev))
            if len(self._tc_tables) > 8:
                self._tc_tables.clear()
            ent = self._tc_tables[key] = (cf, te)
        return ent

    def tc_pack(self):
        """bf16 UMMA weight images for the tensor-core kernels (built on first use)."""
        if self._tc_pack is None:
            nbytes = _lib.lib().gldm_sampler_tc_pack_bytes(ctypes.byref(self.cfg))
            if nbytes < 0:
                raise NotImplementedError(f"precision='bf16': {_lib.lib().gldm_last_error().decode()}")
            with torch.cuda.device(self.device):
                pack = torch.empty(nbytes + 1024, device=self.device, dtype=torch.uint8)
                off = (-pack.data_ptr()) % 1024
                pack = pack[off:off + nbytes]
                _lib.call("gldm_sampler_tc_prepare", ctypes.byref(self.cfg), self.raw.data_ptr(), pack.data_ptr(),
                          _stream(self.device))
            self._tc_pack = pack
        return self._tc_pack


def packed_resnet(module, L, cond_ch):
    cache = module.__dict__.setdefault("_gldm_packs", {})
    sig = _signature(module)
    ent = cache.get((L, cond_ch))
    if ent is None or ent[0] != sig:
        ent = (sig, PackedResNet(module, L, cond_ch))
        cache[(L, cond_ch)] = ent
    return ent[1]


def _check_precision(precision):
    if precision not in PRECISIONS:
        raise ValueError(f"precision must be one of {PRECISIONS}, got {precision!r}")


def _cond3d(z):
    """Conditioning as [B, C, Dc].  A 2-D latent [B, Dc] (PVCNNEncoder with out_channels=1 squeezes its output,
    pc_encoders.py:113-114) is the single-channel case of the multi-channel FiLM (resnets.py:163-175)."""
    if z.ndim == 2:
        return z.unsqueeze(1)
    if z.ndim != 3:
        raise RuntimeError(f"conditioning latent must be [B, Dc] or [B, C, Dc], got {tuple(z.shape)}")
    return z


def class_embedding(module, cls_cond, rows, dev):
    """cls_embed of a class-conditioned denoiser (class_conditioned_resnet.py:43-46, 96-98): cls_cond [rows] / [rows,1] ->
    f32 [rows, emb] = SiLU(Linear(1 -> emb)), added to the time embedding inside the kernels."""
    lin = module.cls_embed[0]
    c = cls_cond.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
    if c.numel() != rows:
        raise RuntimeError(f"class conditioning has {c.numel()} entries for {rows} conditioning rows")
    out = torch.empty((rows, lin.out_features), device=dev, dtype=torch.float32)
    _lib.call("gldm_class_embed", lin.weight.detach().float().contiguous().data_ptr(), lin.bias.detach().float().contiguous().data_ptr(),
              c.data_ptr(), rows, lin.out_features, out.data_ptr(), _stream(dev))
    return out


def resnet_forward(module, x, time, z_cond, precision="fp32", cls_cond=None):
    """One network evaluation: x [B,1,L], time int[B] or None, z_cond [B,C,Dc] -> [B,1,L]."""
    _check_precision(precision)
    _require_cuda(x, "x")
    _require_cuda(z_cond, "z_cond")
    z_cond = _cond3d(z_cond)
    B, _, L = x.shape
    pk = packed_resnet(module, L, z_cond.shape[1])
    dev = x.device
    if cls_cond is not None:
        with torch.cuda.device(dev):
            xin = x.reshape(B, L).contiguous().float()
            zc = z_cond.contiguous().float()
            out = torch.empty((B, L), device=dev, dtype=torch.float32)
            ftime = torch.is_floating_point(time)
            t32 = time.to(device=dev, dtype=torch.float32 if ftime else torch.int32).contiguous()
            ce = class_embedding(module, cls_cond, B, dev)
            ti, tf = (None, t32.data_ptr()) if ftime else (t32.data_ptr(), None)
            if precision == "bf16":
                _lib.call("gldm_denoiser_forward_ex_tc", ctypes.byref(pk.cfg), pk.raw.data_ptr(), pk.tc_pack().data_ptr(),
                          xin.data_ptr(), ti, tf, zc.data_ptr(), ce.data_ptr(), B, out.data_ptr(), _stream(dev))
            else:
                _lib.call("gldm_denoiser_forward_ex_f32", ctypes.byref(pk.cfg), pk.prepared.data_ptr(), xin.data_ptr(), ti, tf,
                          zc.data_ptr(), ce.data_ptr(), B, out.data_ptr(), _stream(dev))
        return out.view(B, 1, L)
    with torch.cuda.device(dev):
        xin = x.reshape(B, L).contiguous().float()
        zc = z_cond.contiguous().float()
        out = torch.empty((B, L), device=dev, dtype=torch.float32)
        ftime = time is not None and torch.is_floating_point(time)      # continuous time (elucidated sampler)
        t32 = None if time is None else time.to(device=dev, dtype=torch.float32 if ftime else torch.int32).contiguous()
        sfx = "_ftime" if ftime else ""
        if precision == "bf16":
            _lib.call("gldm_denoiser_forward_tc" + sfx, ctypes.byref(pk.cfg), pk.raw.data_ptr(), pk.tc_pack().data_ptr(),
                      xin.data_ptr(), t32.data_ptr() if t32 is not None else None, zc.data_ptr(), B, out.data_ptr(),
                      _stream(dev))
        else:
            _lib.call("gldm_denoiser_forward_f32" + sfx, ctypes.byref(pk.cfg), pk.prepared.data_ptr(), xin.data_ptr(),
                      t32.data_ptr() if t32 is not None else None, zc.data_ptr(), B, out.data_ptr(), _stream(dev))
    return out.view(B, 1, L)


SCHED_EDM = 2


def _per_object_class(cls_cond, n, n_obj, gpo):
    """class conditioning per conditioning row: [n_obj] as is; [n] (the reference's per-sample layout) must be constant
    over the grasps of an object"""
    c = cls_cond.reshape(-1)
    if c.numel() == n_obj:
        return c
    if c.numel() == n:
        c2 = c.reshape(n_obj, gpo)
        if not bool((c2 == c2[:, :1]).all()):
            raise NotImplementedError("class conditioning must be the same for all grasps of an object")
        return c2[:, 0].contiguous()
    raise RuntimeError(f"class conditioning has {c.numel()} entries for {n} samples of {n_obj} objects")


def sampler_program(denoiser, x_init, z_obj, grasps_per_obj, times, rows, n_out_slots, clip_sample=False, noise=None, seed=0,
                    return_all=False, precision="fp32", cls_cond=None, sched_kind=SCHED_EDM):
    """Persistent sampler on an evaluation program (the elucidated samplers: one launch for all network evaluations and
    updates).  times f32[n_evals] (c_noise of every evaluation), rows f32[n_evals,16] (include/graspldm_b200.h,
    GldmSamplerArgs) - both host tensors.  Returns (x [n,1,D], x_all [n_out_slots,n,1,D] or None)."""
    _check_precision(precision)
    _require_cuda(x_init, "x_init")
    _require_cuda(z_obj, "z_cond")
    n, _, D = x_init.shape
    z_obj = _cond3d(z_obj)
    pk = packed_resnet(denoiser, D, z_obj.shape[1])
    dev = x_init.device
    n_evals = rows.shape[0]
    with torch.cuda.device(dev):
        xin = x_init.reshape(n, D).contiguous().float()
        zc = z_obj.contiguous().float()
        out = torch.empty((n, D), device=dev, dtype=torch.float32)
        x_all = torch.empty((n_out_slots, n, D), device=dev, dtype=torch.float32) if return_all else None
        nz = None
        if noise is not None:
            _require_cuda(noise, "noise")
            nz = noise.reshape(-1, n, D).contiguous().float()
        cf = rows.detach().to(device=dev, dtype=torch.float32).contiguous()
        tm = times.detach().to(device=dev, dtype=torch.float32).contiguous()
        a = GldmSamplerArgs()
        a.x_init, a.z_obj, a.n, a.grasps_per_obj = xin.data_ptr(), zc.data_ptr(), n, int(grasps_per_obj)
        a.sched_kind, a.n_steps, a.coef, a.times = int(sched_kind), n_evals, cf.data_ptr(), tm.data_ptr()
        a.clip_sample, a.noise, a.seed = int(bool(clip_sample)), (nz.data_ptr() if nz is not None else None), int(seed) & (2 ** 64 - 1)
        a.x_out, a.x_all = out.data_ptr(), (x_all.data_ptr() if x_all is not None else None)
        ce = None
        if cls_cond is not None:
            ce = class_embedding(denoiser, _per_object_class(cls_cond, n, zc.shape[0], int(grasps_per_obj)), zc.shape[0], dev)
            a.cls_emb = ce.data_ptr()
        if precision == "bf16":
            te = torch.empty((n_evals, pk.cfg.emb_dim), device=dev, dtype=torch.float32)
            _lib.call("gldm_time_embed_table_f", ctypes.byref(pk.cfg), pk.raw.data_ptr(), tm.data_ptr(), n_evals, te.data_ptr(), _stream(dev))
            a.te = te.data_ptr()
            pack = pk.tc_pack()
            tok = SECTIONS.start("sampler", dev)
            _lib.call("gldm_sampler_run_ex_tc", ctypes.byref(pk.cfg), pk.raw.data_ptr(), pack.data_ptr(), ctypes.byref(a), _stream(dev))
        else:
            tok = SECTIONS.start("sampler", dev)
            _lib.call("gldm_sampler_run_ex_f32", ctypes.byref(pk.cfg), pk.prepared.data_ptr(), ctypes.byref(a), _stream(dev))
        SECTIONS.stop(tok)
    return out.view(n, 1, D), (x_all.view(n_out_slots, n, 1, D) if x_all is not None else None)


def sampler_run(denoiser, x_T, z_obj, grasps_per_obj, timesteps, coef, sched_kind, clip_sample, noise=None,
                seed=0, return_all=False, precision="fp32", cls_cond=None):
    """Whole reverse-diffusion loop in one launch.  x_T [n,1,D]; z_obj [n_obj,C,Dc]; timesteps list[int];
    coef float32 [n_steps,8] (host).  Returns (x_0 [n,1,D], x_all [n_steps+1,n,1,D] or None)."""
    _check_precision(precision)
    if cls_cond is not None:
        return _sampler_run_class_conditioned(denoiser, x_T, z_obj, grasps_per_obj, timesteps, coef, sched_kind, clip_sample,
                                              noise, seed, return_all, precision, cls_cond)
    _require_cuda(x_T, "x_T")
    _require_cuda(z_obj, "z_cond")
    n, _, D = x_T.shape
    z_obj = _cond3d(z_obj)
    pk = packed_resnet(denoiser, D, z_obj.shape[1])
    dev = x_T.device
    n_steps = len(timesteps)
    ts = (ctypes.c_int * n_steps)(*[int(t) for t in timesteps])
    cf = coef.detach().cpu().contiguous().float()
    assert cf.shape == (n_steps, 8)
    with torch.cuda.device(dev):
        xin = x_T.reshape(n, D).contiguous().float()
        zc = z_obj.contiguous().float()
        out = torch.empty((n, D), device=dev, dtype=torch.float32)
        x_all = torch.empty((n_steps + 1, n, D), device=dev, dtype=torch.float32) if return_all else None
        nz = None
        if noise is not None:
            _require_cuda(noise, "noise")
            nz = noise.reshape(n_steps, n, D).contiguous().float()
        tail = (n, int(grasps_per_obj), n_steps, ctypes.cast(ts, ctypes.c_void_p), cf.data_ptr(), int(sched_kind),
                int(bool(clip_sample)), nz.data_ptr() if nz is not None else None, int(seed) & (2 ** 64 - 1),
                out.data_ptr(), x_all.data_ptr() if x_all is not None else None, _stream(dev))
        if precision == "bf16":                                   # built / cached before the timed section
            pack = pk.tc_pack()
            cf_dev, te_dev = pk.tc_tables(timesteps, coef)
        tok = SECTIONS.start("sampler", dev)
        if precision == "bf16":
            _lib.call("gldm_sampler_run_tc_dev", ctypes.byref(pk.cfg), pk.raw.data_ptr(), pack.data_ptr(), xin.data_ptr(),
                      zc.data_ptr(), n, int(grasps_per_obj), n_steps, cf_dev.data_ptr(), te_dev.data_ptr(), int(sched_kind),
                      int(bool(clip_sample)), nz.data_ptr() if nz is not None else None, int(seed) & (2 ** 64 - 1),
                      out.data_ptr(), x_all.data_ptr() if x_all is not None else None, _stream(dev))
        else:
            _lib.call("gldm_sampler_run_f32", ctypes.byref(pk.cfg), pk.prepared.data_ptr(), xin.data_ptr(),
                      zc.data_ptr(), *tail)
        SECTIONS.stop(tok)
    return out.view(n, 1, D), (x_all.view(n_steps + 1, n, 1, D) if x_all is not None else None)


def _sampler_run_class_conditioned(denoiser, x_T, z_obj, grasps_per_obj, timesteps, coef, sched_kind, clip_sample, noise, seed,
                                   return_all, precision, cls_cond):
    """DDPM / DDIM loop of a ClassTimeConditionedResNet1D: the class embedding of every object rides along
    (GldmSamplerArgs.cls_emb) and is added to the time embedding of every step inside the kernel."""
    _require_cuda(x_T, "x_T")
    _require_cuda(z_obj, "z_cond")
    n, _, D = x_T.shape
    z_obj = _cond3d(z_obj)
    pk = packed_resnet(denoiser, D, z_obj.shape[1])
    dev = x_T.device
    n_steps = len(timesteps)
    with torch.cuda.device(dev):
        xin = x_T.reshape(n, D).contiguous().float()
        zc = z_obj.contiguous().float()
        out = torch.empty((n, D), device=dev, dtype=torch.float32)
        x_all = torch.empty((n_steps + 1, n, D), device=dev, dtype=torch.float32) if return_all else None
        nz = noise.reshape(n_steps, n, D).contiguous().float() if noise is not None else None
        ts = torch.tensor([int(t) for t in timesteps], dtype=torch.int32, device=dev)
        cf = coef.detach().to(device=dev, dtype=torch.float32).contiguous()
        ce = class_embedding(denoiser, _per_object_class(cls_cond, n, zc.shape[0], int(grasps_per_obj)), zc.shape[0], dev)
        a = GldmSamplerArgs()
        a.x_init, a.z_obj, a.n, a.grasps_per_obj = xin.data_ptr(), zc.data_ptr(), n, int(grasps_per_obj)
        a.sched_kind, a.n_steps, a.coef, a.timesteps = int(sched_kind), n_steps, cf.data_ptr(), ts.data_ptr()
        a.clip_sample, a.noise, a.seed = int(bool(clip_sample)), (nz.data_ptr() if nz is not None else None), int(seed) & (2 ** 64 - 1)
        a.cls_emb, a.x_out, a.x_all = ce.data_ptr(), out.data_ptr(), (x_all.data_ptr() if x_all is not None else None)
        if precision == "bf16":
            _, te = pk.tc_tables(timesteps, coef)
            a.te = te.data_ptr()
            _lib.call("gldm_sampler_run_ex_tc", ctypes.byref(pk.cfg), pk.raw.data_ptr(), pk.tc_pack().data_ptr(), ctypes.byref(a), _stream(dev))
        else:
            _lib.call("gldm_sampler_run_ex_f32", ctypes.byref(pk.cfg), pk.prepared.data_ptr(), ctypes.byref(a), _stream(dev))
    return out.view(n, 1, D), (x_all.view(n_steps + 1, n, 1, D) if x_all is not None else None)


def decoder_forward(decoder, z_h, z_obj, grasps_per_obj, precision="fp32"):
    """ConditionalGraspPoseDecoder: z_h [n,D], z_obj [n_obj,C,Dc] -> (tmrp [n,6], logits [n,1])."""
    _check_precision(precision)
    _require_cuda(z_h, "z_h")
    _require_cuda(z_obj, "cond")
    n, D = z_h.shape
    z_obj = _cond3d(z_obj)
    net = decoder.net
    L = decoder.feature_resolution
    pk = packed_resnet(net, L, z_obj.shape[1])
    dev = z_h.device
    cache = decoder.__dict__.setdefault("_gldm_head", {})
    sig = _signature(decoder)
    if cache.get("sig") != sig:
        cache["head"] = _flatten_nopad([decoder.in_layer.weight, decoder.in_layer.bias, decoder.tmrp.weight,
                                        decoder.tmrp.bias, decoder.class_logits.weight, decoder.class_logits.bias], dev)
        cache["sig"] = sig
    head = cache["head"]
    with torch.cuda.device(dev):
        zin = z_h.contiguous().float()
        zc = z_obj.contiguous().float()
        tmrp = torch.empty((n, 6), device=dev, dtype=torch.float32)
        logit = torch.empty((n, 1), device=dev, dtype=torch.float32)
        pack = pk.tc_pack() if precision == "bf16" else None      # built before the timed section
        tok = SECTIONS.start("decoder", dev)
        if precision == "bf16":
            _lib.call("gldm_decoder_forward_tc", ctypes.byref(pk.cfg), pk.raw.data_ptr(), pack.data_ptr(), head.data_ptr(),
                      D, zin.data_ptr(), zc.data_ptr(), n, int(grasps_per_obj), tmrp.data_ptr(), logit.data_ptr(),
                      _stream(dev))
        else:
            _lib.call("gldm_decoder_forward_f32", ctypes.byref(pk.cfg), pk.prepared.data_ptr(), head.data_ptr(), D,
                      zin.data_ptr(), zc.data_ptr(), n, int(grasps_per_obj), tmrp.data_ptr(), logit.data_ptr(),
                      _stream(dev))
        SECTIONS.stop(tok)
    return tmrp, logit


def _flatten_nopad(tensors, device):
    return torch.cat([t.detach().to(device=device, dtype=torch.float32).reshape(-1) for t in tensors]).contiguous()


def pose_postprocess(tmrp, logit, grasp_mean, grasp_std, grasps_per_obj=None):
    """tmrp [n,6], logit [n,1] -> (grasp_tmrp [n,6], H [n,4,4], confidence [n,1]).
    grasp_mean / grasp_std: one shared row ([6] or [1,6]) or one row per object ([n / grasps_per_obj, 6])."""
    _require_cuda(tmrp, "tmrp")
    n = tmrp.shape[0]
    dev = tmrp.device
    with torch.cuda.device(dev):
        t = tmrp.contiguous().float()
        lg = logit.contiguous().float()
        gm = grasp_mean.to(dev).reshape(-1, 6).contiguous().float()
        gs = grasp_std.to(dev).reshape(-1, 6).contiguous().float()
        gpo = int(grasps_per_obj) if grasps_per_obj else max(n, 1)
        gt = torch.empty((n, 6), device=dev, dtype=torch.float32)
        H = torch.empty((n, 4, 4), device=dev, dtype=torch.float32)
        conf = torch.empty((n, 1), device=dev, dtype=torch.float32)
        _lib.call("gldm_pose_postprocess_rows", t.data_ptr(), lg.data_ptr(), gm.data_ptr(), gs.data_ptr(), n, gpo,
                  gm.shape[0], gs.shape[0], gt.data_ptr(), H.data_ptr(), conf.data_ptr(), _stream(dev))
    return gt, H, conf


def normalize_clouds(pc, pc_shift, pc_scale, grasp_shift):
    """pc [b,n,3] raw -> (pc_norm [b,n,3], pc_mean [b,3], grasp_mean [b,6]); inference_base.py:182-212."""
    _require_cuda(pc, "pc")
    dev = pc.device
    b, n, _ = pc.shape
    with torch.cuda.device(dev):
        src = pc.contiguous().float()
        f = lambda v: v.to(dev).reshape(-1).contiguous().float()
        ps, pscale, gsh = f(pc_shift), f(pc_scale), f(grasp_shift)
        assert ps.numel() == 3 and pscale.numel() == 3 and gsh.numel() == 6
        out = torch.empty_like(src)
        pm = torch.empty((b, 3), device=dev, dtype=torch.float32)
        gm = torch.empty((b, 6), device=dev, dtype=torch.float32)
        _lib.call("gldm_normalize_clouds", src.data_ptr(), ps.data_ptr(), pscale.data_ptr(), gsh.data_ptr(), b, n,
                  out.data_ptr(), pm.data_ptr(), gm.data_ptr(), _stream(dev))
    return out, pm, gm


def _conv3d(x, w_f32, w_img, bias, B, ci, co, r, y, st):
    """Conv3d k3 p1: tensor-core kernel when a packed bf16 weight image is given, strict-fp32 SIMT kernel otherwise."""
    if w_img is None:
        _lib.call("gldm_conv3d_k3_f32", x.data_ptr(), w_f32.data_ptr(), bias.data_ptr(), B, ci, co, r, y.data_ptr(), st)
        return
    scratch = _aligned_bytes(_lib.lib().gldm_conv3d_tc_grid_bytes(B, ci, r), x.device, 256)
    _lib.call("gldm_conv3d_k3_tc", x.data_ptr(), w_img.data_ptr(), bias.data_ptr(), B, ci, co, r, scratch.data_ptr(),
              y.data_ptr(), st)


def _pointwise_tail_tc(pk, feats, st):
    """SharedMLPs after the last PVConv -> conv_downscale -> out_layer.0 on the tensor cores.  Activations stay in
    HBM as bf16 UMMA images (rows = points of all clouds); only the [B, C_out, N] result is fp32."""
    dev = feats.device
    B, C, N = feats.shape
    rows = B * N
    if rows % 128:
        raise NotImplementedError("precision='bf16': clouds * points must be a multiple of 128")
    L = _lib.lib()
    img = _aligned_bytes(L.gldm_gemm_tc_image_bytes(rows, C), dev)
    _lib.call("gldm_gemm_tc_to_image", feats.data_ptr(), B, C, N, img.data_ptr(), st)
    layers = pk.tc_weights()
    proj = pk.tc_projection()
    if proj is not None:
        # conv_downscale and out_layer.0 are composed into one [C_out, width] projection that the epilogue of the last
        # SharedMLP applies to its fp32 activations: neither the widest activation nor conv_downscale's output exists
        for layer in layers[:-2]:
            out = _aligned_bytes(L.gldm_gemm_tc_image_bytes(rows, layer["n"]), dev)
            _lib.call("gldm_gemm_tc_run", img.data_ptr(), layer["img"].data_ptr(), layer["scale"].data_ptr(),
                      layer["shift"].data_ptr(), rows, layer["k"], layer["n"], layer["relu"], out.data_ptr(), st)
            img = out
        layer = layers[-2]
        part = torch.empty((layer["n"] // 128, rows, 4), device=dev, dtype=torch.float32)
        h = torch.empty((B, pk.out_channels, N), device=dev, dtype=torch.float32)
        _lib.call("gldm_gemm_tc_run_proj", img.data_ptr(), layer["img"].data_ptr(), layer["scale"].data_ptr(),
                  layer["shift"].data_ptr(), rows, layer["k"], layer["n"], layer["relu"], proj[0].data_ptr(),
                  proj[1].data_ptr(), pk.out_channels, N, part.data_ptr(), h.data_ptr(), st)
        return h
    for layer in layers:
        out = _aligned_bytes(L.gldm_gemm_tc_image_bytes(rows, layer["n"]), dev)
        _lib.call("gldm_gemm_tc_run", img.data_ptr(), layer["img"].data_ptr(),
                  layer["scale"].data_ptr() if layer["scale"] is not None else None,
                  layer["shift"].data_ptr() if layer["shift"] is not None else None,
                  rows, layer["k"], layer["n"], layer["relu"], out.data_ptr(), st)
        img, k = out, layer["n"]
    h = torch.empty((B, pk.out_channels, N), device=dev, dtype=torch.float32)
    _lib.call("gldm_gemm_tc_image_small_co", img.data_ptr(), pk.wo.data_ptr(), pk.bo.data_ptr(), rows, k,
              pk.out_channels, N, h.data_ptr(), st)
    return h


# --------------------------------------------------------------------------------------------------
# PVCNN encoder (strict fp32 path)
# --------------------------------------------------------------------------------------------------
def _fold_bn(conv, bn):
    """Conv bias + eval-mode BatchNorm as y = scale * (W x) + shift (one-time constant folding)."""
    scale = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)
    shift = (conv.bias.detach() - bn.running_mean.detach()) * scale + bn.bias.detach()
    return scale.float().contiguous(), shift.float().contiguous()


def pack_point_block(blk):
    """Kernel-ready constants of one PVCNN point block: a PVConv (pvconv.py:13-84) or a single-layer SharedMLP."""
    if hasattr(blk, "voxel_layers"):
        vl = [m for m in blk.voxel_layers if not isinstance(m, torch.nn.Dropout)]      # conv, gn, swish, conv, gn, swish, se
        c1, g1, c2, g2, se = vl[0], vl[1], vl[3], vl[4], vl[6]
        pw = blk.point_features.layers
        assert len(pw) == 3, "PVConv's point branch is a single-layer SharedMLP"
        sc, sh = _fold_bn(pw[0], pw[1])
        return dict(
            kind="pvconv", r=blk.resolution, cin=blk.in_channels, cout=blk.out_channels,
            normalize=bool(blk.voxelization.normalize), vox_eps=float(blk.voxelization.eps),
            se_relu=isinstance(se.fc[1], torch.nn.ReLU),
            w1_raw=c1.weight.detach().float().contiguous(), w2_raw=c2.weight.detach().float().contiguous(),
            # Conv3d weights [co,ci,3,3,3] -> [ci,27,co] (layout only)
            w1=c1.weight.detach().permute(1, 2, 3, 4, 0).reshape(c1.in_channels, 27, c1.out_channels).contiguous().float(),
            b1=c1.bias.detach().float().contiguous(), g1w=g1.weight.detach().float().contiguous(),
            g1b=g1.bias.detach().float().contiguous(), eps1=float(g1.eps),
            w2=c2.weight.detach().permute(1, 2, 3, 4, 0).reshape(c2.in_channels, 27, c2.out_channels).contiguous().float(),
            b2=c2.bias.detach().float().contiguous(), g2w=g2.weight.detach().float().contiguous(),
            g2b=g2.bias.detach().float().contiguous(), eps2=float(g2.eps), groups=int(g1.num_groups),
            se1=se.fc[0].weight.detach().float().contiguous(), se2=se.fc[2].weight.detach().float().contiguous(),
            pw=pw[0].weight.detach().reshape(pw[0].out_channels, pw[0].in_channels).float().contiguous(),
            pscale=sc, pshift=sh)
    pw = blk.layers
    assert len(pw) == 3, "single-layer SharedMLP block"
    sc, sh = _fold_bn(pw[0], pw[1])
    return dict(kind="mlp", cin=pw[0].in_channels, cout=pw[0].out_channels,
                pw=pw[0].weight.detach().reshape(pw[0].out_channels, pw[0].in_channels).float().contiguous(),
                pscale=sc, pshift=sh)


def pvconv_forward_f32(blk, feats, coords):
    """PVConv.forward on the strict-fp32 kernels (pvconv.py:76-84): voxelize -> Conv3d, GroupNorm, Swish (x2) -> SE ->
    trilinear devoxelize + point branch.  feats [B,Cin,N], coords [B,3,N] -> [B,Cout,N].  `normalize=True`
    (voxelization.py:19-30, PVCNN2's blocks) takes the per-cloud scaling through a few element-wise device operations and
    the avg_voxelize operator; `normalize=False` (PVCNN) is one fused kernel."""
    dev = feats.device
    st = _stream(dev)
    B, ci, N = feats.shape
    r, co = blk["r"], blk["cout"]
    r3 = r ** 3
    with torch.cuda.device(dev):
        feats = feats.contiguous().float()
        coords = coords.contiguous().float()
        if blk["normalize"]:
            nc = coords - coords.mean(2, keepdim=True)
            nc = nc / (nc.norm(dim=1, keepdim=True).max(dim=2, keepdim=True).values * 2.0 + blk["vox_eps"]) + 0.5
            norm = torch.clamp(nc * r, 0, r - 1).contiguous()
            vox = torch.round(norm).to(torch.int32).contiguous()
            grid = torch.empty((B, ci, r3), device=dev, dtype=torch.float32)
            ind = torch.empty((B, N), device=dev, dtype=torch.int32)
            cnt = torch.empty((B, r3), device=dev, dtype=torch.int32)
            _lib.call("gldm_avg_voxelize_forward", feats.data_ptr(), vox.data_ptr(), B, ci, N, r, grid.data_ptr(),
                      ind.data_ptr(), cnt.data_ptr(), st)
        else:
            grid = torch.empty((B, ci, r3), device=dev, dtype=torch.float32)
            norm = torch.empty((B, 3, N), device=dev, dtype=torch.float32)
            _lib.call("gldm_voxelize_fused", feats.data_ptr(), coords.data_ptr(), B, ci, N, r, grid.data_ptr(),
                      norm.data_ptr(), None, st)
        y1 = torch.empty((B, co, r3), device=dev, dtype=torch.float32)
        _lib.call("gldm_conv3d_k3_f32", grid.data_ptr(), blk["w1"].data_ptr(), blk["b1"].data_ptr(), B, ci, co, r, y1.data_ptr(), st)
        _lib.call("gldm_groupnorm_swish_f32", y1.data_ptr(), blk["g1w"].data_ptr(), blk["g1b"].data_ptr(), B, co, r3,
                  blk["groups"], blk["eps1"], None, st)
        y2 = torch.empty((B, co, r3), device=dev, dtype=torch.float32)
        _lib.call("gldm_conv3d_k3_f32", y1.data_ptr(), blk["w2"].data_ptr(), blk["b2"].data_ptr(), B, co, co, r, y2.data_ptr(), st)
        se_mean = torch.empty((B, co), device=dev, dtype=torch.float32)
        _lib.call("gldm_groupnorm_swish_f32", y2.data_ptr(), blk["g2w"].data_ptr(), blk["g2b"].data_ptr(), B, co, r3,
                  blk["groups"], blk["eps2"], se_mean.data_ptr(), st)
        gate = torch.empty((B, co), device=dev, dtype=torch.float32)
        _lib.call("gldm_se_gate_relu_f32" if blk["se_relu"] else "gldm_se_gate_f32", se_mean.data_ptr(), blk["se1"].data_ptr(),
                  blk["se2"].data_ptr(), B, co, blk["se1"].shape[0], gate.data_ptr(), st)
        pt = _pw(feats, blk["pw"], blk["pscale"], blk["pshift"], None, 1)
        fused = torch.empty((B, co, N), device=dev, dtype=torch.float32)
        _lib.call("gldm_devox_gate_add_f32", norm.data_ptr(), y2.data_ptr(), gate.data_ptr(), pt.data_ptr(), B, co, N, r,
                  fused.data_ptr(), st)
    return fused


def packed_block(module):
    """pack_point_block cached on the module (rebuilt when its parameters change)."""
    sig = _signature(module)
    ent = module.__dict__.get("_gldm_block")
    if ent is None or ent[0] != sig:
        _require_cuda(next(module.parameters()), "model parameters")
        ent = (sig, pack_point_block(module))
        module.__dict__["_gldm_block"] = ent
    return ent[1]


def sa_group_mlp_max(mlp, coords, centers, feats, idx, include_coords=True):
    """Fused grouping -> SharedMLP(dim=2) -> max of one PointNetSAModule branch (csrc/set_abstraction.cu).
    coords [B,3,N], centers [B,3,M], feats [B,C,N] or None, idx i32 [B,M,U] -> [B,C_out,M]."""
    _require_cuda(coords, "coords")
    dev = coords.device
    sig = _signature(mlp)
    ent = mlp.__dict__.get("_gldm_sa")
    if ent is None or ent[0] != sig:
        layers = []
        for i in range(0, len(mlp.layers), 3):
            conv, bn = mlp.layers[i], mlp.layers[i + 1]
            sc, sh = _fold_bn(conv, bn)
            wt = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).t().contiguous().float()
            layers.append((wt, sc, sh, conv.out_channels))
        ent = (sig, layers)
        mlp.__dict__["_gldm_sa"] = ent
    layers = ent[1]
    B, _, N = coords.shape
    M, U = idx.shape[1], idx.shape[2]
    C = 0 if feats is None else feats.shape[1]
    nl = len(layers)
    widths = (ctypes.c_int * nl)(*[l[3] for l in layers])
    arr = lambda k: (ctypes.c_void_p * nl)(*[l[k].data_ptr() for l in layers])
    with torch.cuda.device(dev):
        out = torch.empty((B, layers[-1][3], M), device=dev, dtype=torch.float32)
        f = None if feats is None else feats.contiguous().float()
        _lib.call("gldm_sa_mlp_max_f32", coords.contiguous().data_ptr(), centers.contiguous().data_ptr(),
                  f.data_ptr() if f is not None else None, idx.contiguous().data_ptr(), B, C, N, M, U, int(include_coords), nl,
                  ctypes.cast(widths, ctypes.c_void_p), ctypes.cast(arr(0), ctypes.c_void_p), ctypes.cast(arr(1), ctypes.c_void_p),
                  ctypes.cast(arr(2), ctypes.c_void_p), out.data_ptr(), _stream(dev))
    return out


class PackedEncoder:
    """Device-resident, kernel-ready weights of a PVCNNEncoder."""

    def __init__(self, enc):
        p0 = next(enc.parameters())
        _require_cuda(p0, "model parameters")
        self.device = p0.device
        self.blocks = [pack_point_block(blk) for blk in enc.pvcnn_modules.point_features]
        cd = enc.conv_downscale
        self.wd = cd.weight.detach().reshape(cd.out_channels, cd.in_channels).float().contiguous()
        self.bd = cd.bias.detach().float().contiguous()
        co = enc.out_layer[0]
        self.wo = co.weight.detach().reshape(co.out_channels, co.in_channels).float().contiguous()
        self.bo = co.bias.detach().float().contiguous()
        self.wl = enc.out_layer[1].weight.detach().float().contiguous()
        self.bl = enc.out_layer[1].bias.detach().float().contiguous()
        self.out_channels = co.out_channels
        self.out_features = enc.out_layer[1].out_features
        self._tc = None
        self._tc_conv = None
        self._proj = None

    def tc_conv_weights(self):
        """bf16 UMMA images of the Conv3d weights that qualify for the tensor-core kernel (16 <= ci, co <= 128);
        one (img1 | None, img2 | None) pair per PVConv block."""
        if self._tc_conv is None:
            dev, out = self.device, []
            with torch.cuda.device(dev):
                for blk in self.blocks:
                    if blk["kind"] != "pvconv":
                        out.append(None)
                        continue
                    pair = []
                    for w in (blk["w1_raw"], blk["w2_raw"]):
                        co, ci = w.shape[0], w.shape[1]
                        if 16 <= ci and co <= 128:
                            img = _aligned_bytes(_lib.lib().gldm_conv3d_tc_weight_bytes(ci), dev)
                            _lib.call("gldm_conv3d_tc_pack_weight", w.data_ptr(), co, ci, img.data_ptr(), _stream(dev))
                            pair.append(img)
                        elif ci < 16 and co <= 128 and co % 8 == 0 and os.environ.get("GLDM_CONV3D_TC16", "1") != "0":
                            # narrow first layer (3 -> 48): K = 16 image, tagged so that the voxel branch picks the kernel
                            img = _aligned_bytes(_lib.lib().gldm_conv3d_tc16_weight_bytes(), dev)
                            _lib.call("gldm_conv3d_tc16_pack_weight", w.data_ptr(), co, ci, img.data_ptr(), _stream(dev))
                            img.narrow_k = True
                            pair.append(img)
                        else:
                            pair.append(None)
                    out.append(pair)
            self._tc_conv = out
        return self._tc_conv

    def tc_projection(self):
        """(proj_w [width, 4], proj_bias [C_out]) of conv_downscale followed by out_layer.0, composed in fp64 (two affine
        maps in a row, pc_encoders.py:104-112 without global attention), or None when there is no SharedMLP tail to
        fuse it into, C_out > 4, or GLDM_FOLD_DOWNSCALE=0 (the layer-by-layer tensor-core chain)."""
        if os.environ.get("GLDM_FOLD_DOWNSCALE", "1") == "0" or self.out_channels > 4:
            return None
        if not self.blocks or self.blocks[-1]["kind"] != "mlp" or self.blocks[-1]["pscale"] is None:
            return None
        if self._proj is None:
            w, b = compose_affine(self.wo, self.bo, self.wd, self.bd)     # [C_out, width], [C_out]
            pw = torch.zeros((w.shape[1], 4), device=self.device, dtype=torch.float32)
            pw[:, :w.shape[0]] = w.t()
            self._proj = (pw.contiguous(), b.contiguous())
        return self._proj

    def tc_weights(self):
        """bf16 UMMA images of the point-wise layers that run on the tensor cores (built on first use):
        every SharedMLP after the last PVConv block, then conv_downscale."""
        if self._tc is None:
            dev = self.device
            layers = []
            tail = [b for b in self.blocks if b["kind"] == "mlp"]
            specs = [(b["pw"], b["pscale"], b["pshift"], 1) for b in tail] + [(self.wd, None, self.bd, 0)]
            with torch.cuda.device(dev):
                for w, sc, sh, relu in specs:
                    n_out, k = w.shape
                    if n_out % 128:
                        raise NotImplementedError(f"precision='bf16': point-wise layer width {n_out} is not a multiple of 128")
                    nbytes = _lib.lib().gldm_gemm_tc_image_bytes(n_out, k)
                    img = _aligned_bytes(nbytes, dev)
                    _lib.call("gldm_gemm_tc_pack_weight", w.data_ptr(), n_out, k, img.data_ptr(), _stream(dev))
                    layers.append(dict(img=img, scale=sc, shift=sh, relu=relu, k=k, n=n_out))
            self._tc = layers
        return self._tc


def compose_affine(w2, b2, w1, b1):
    """y = W2 (W1 x + b1) + b2 as one affine map (W, b), composed in fp64 and rounded to fp32 once: conv_downscale followed
    by out_layer.0 (pc_encoders.py:104-112 with use_global_attention=False: no non-linearity in between)."""
    w = w2.double() @ w1.double()
    b = w2.double() @ b1.double() + b2.double()
    return w.float(), b.float()


def _aligned_bytes(nbytes, dev, align=1024):
    buf = torch.empty(nbytes + align, device=dev, dtype=torch.uint8)
    off = (-buf.data_ptr()) % align
    return buf[off:off + nbytes]


def packed_encoder(enc):
    sig = _signature(enc)
    ent = enc.__dict__.get("_gldm_pack")
    if ent is None or ent[0] != sig:
        ent = (sig, PackedEncoder(enc))
        enc.__dict__["_gldm_pack"] = ent
    return ent[1]


def _pw(x, w, scale, shift, add, act):
    b, ci, n = x.shape
    co = w.shape[0]
    y = torch.empty((b, co, n), device=x.device, dtype=torch.float32)
    _lib.call("gldm_pointwise_conv_f32", x.data_ptr(), w.data_ptr(), scale.data_ptr() if scale is not None else None,
              shift.data_ptr() if shift is not None else None, add.data_ptr() if add is not None else None,
              b, ci, co, n, act, y.data_ptr(), _stream(x.device))
    return y


def encoder_forward(enc, xyz, max_clouds_per_pass=256, precision="fp32"):
    """PVCNNEncoder.forward: xyz [B,N,3] -> z_pc [B,C_out,F] (squeezed when C_out == 1)."""
    _require_cuda(xyz, "xyz")
    _check_precision(precision)
    pk = packed_encoder(enc)
    if precision == "bf16":
        pk.tc_weights()
        pk.tc_conv_weights()
    outs = []
    tok = SECTIONS.start("encoder", xyz.device)
    for s in range(0, xyz.shape[0], max_clouds_per_pass):
        outs.append(_encoder_pass(pk, xyz[s:s + max_clouds_per_pass], precision))
    SECTIONS.stop(tok)
    out = torch.cat(outs) if len(outs) > 1 else outs[0]
    return out.squeeze(1) if out.shape[-2] == 1 else out


_CL_CACHE_MAX = 64      # buffers kept per encoder (tag x stream); least recently used ones are dropped beyond that


def _cl_grid(pk, dev, tag, rows, stride, dtype):
    """Zero-initialised padded channels-last grid, cached per (tag, stream): the kernels never write halo rows or padding
    channels with anything but zero, so the zeros survive from call to call; the stream key keeps concurrently running
    passes (bench --streams) apart.  One buffer per key, sized for the LARGEST batch seen: a smaller batch uses a prefix
    (its halo rows are the same rows, still zero), a larger one replaces the buffer.  Bounded (LRU) - see clear_cl_cache."""
    cache = pk.__dict__.setdefault("_cl_cache", {})
    key = (tag, torch.cuda.current_stream(dev).cuda_stream, dev.index, stride if rows > 1 else 0, dtype)
    need = rows * stride
    buf = cache.pop(key, None)
    if buf is None or buf.numel() < need:
        buf = torch.zeros(need, device=dev, dtype=dtype)
        assert buf.data_ptr() % 256 == 0
    cache[key] = buf                              # (re)inserted last = most recently used
    while len(cache) > _CL_CACHE_MAX:
        cache.pop(next(iter(cache)))
    return buf[:need].view(rows, stride)


def clear_cl_cache(enc):
    """Release the cached channels-last grids of an encoder (they are re-created, zeroed, on the next bf16 pass)."""
    ent = enc.__dict__.get("_gldm_pack")
    if ent is not None:
        ent[1].__dict__.pop("_cl_cache", None)


def _pvconv_voxel_branch_tc(pk, bi, blk, feats, coords, B, N, st, dev):
    """Voxel branch of one PVConv on the fused channels-last path: voxelize -> Conv3d -> GN+Swish -> Conv3d -> GN+Swish
    (+SE squeeze) -> SE gate -> devoxelize(+ point branch).  Returns the fused point features [B, co, N]."""
    r, ci, co = blk["r"], blk["cin"], blk["cout"]
    assert blk["groups"] == 8 and co % 8 == 0
    w1_img, w2_img = pk.tc_conv_weights()[bi]
    r3, P = r ** 3, (r + 2) ** 3
    rows = B * P
    norm = torch.empty((B, 3, N), device=dev, dtype=torch.float32)
    # tensor-core first conv: the voxelisation writes the Conv3d operand (zero-padded channels-last bf16 rows) directly;
    # otherwise the reference's fp32 [B, C, r^3] grid
    direct = w1_img is not None and (ci <= 4 or ci % 8 == 0) and os.environ.get("GLDM_VOX_CL") != "0"
    grid = None
    if not direct:
        grid = torch.empty((B, ci, r3), device=dev, dtype=torch.float32)
        _lib.call("gldm_voxelize_fused", feats.data_ptr(), coords.data_ptr(), B, ci, N, r, grid.data_ptr(), norm.data_ptr(),
                  None, st)
    stats = torch.empty((2, B, 8, 2), device=dev, dtype=torch.float64)
    se_sum = torch.empty((B, co), device=dev, dtype=torch.float64)
    # workspace of the bit-reproducible (atomic-free) statistics: per-block partials, added in a fixed order
    ws = _cl_grid(pk, dev, ("ws", bi), 1, (_lib.lib().gldm_voxel_ws_bytes(B, max(ci, co), r) + 7) // 8, torch.float64)
    cpad_o = -(-co // 64) * 64
    if w1_img is not None and getattr(w1_img, "narrow_k", False):
        x16 = _cl_grid(pk, dev, ("x16", bi), rows, 16, torch.bfloat16)
        y1 = _cl_grid(pk, dev, ("y1", bi), rows, cpad_o, torch.bfloat16)
        if direct:
            _lib.call("gldm_voxelize_fused_cl", feats.data_ptr(), coords.data_ptr(), B, ci, N, r, x16.data_ptr(), 16,
                      norm.data_ptr(), st)
        _lib.call("gldm_conv3d_tc16_cl", None if direct else grid.data_ptr(), w1_img.data_ptr(), blk["b1"].data_ptr(), B, ci, co, r,
                  x16.data_ptr(), y1.data_ptr(), cpad_o, stats[0].data_ptr(), ws.data_ptr(), st)
        _lib.call("gldm_gn_swish_cl", y1.data_ptr(), 0, cpad_o, stats[0].data_ptr(), blk["g1w"].data_ptr(),
                  blk["g1b"].data_ptr(), B, co, r, blk["eps1"], None, None, st)
    elif w1_img is not None:
        x_cl = _cl_grid(pk, dev, ("x", bi), rows, -(-ci // 64) * 64, torch.bfloat16)
        if direct:
            _lib.call("gldm_voxelize_fused_cl", feats.data_ptr(), coords.data_ptr(), B, ci, N, r, x_cl.data_ptr(),
                      -(-ci // 64) * 64, norm.data_ptr(), st)
        else:
            _lib.call("gldm_cl_pad", grid.data_ptr(), B, ci, r, x_cl.data_ptr(), st)
        y1 = _cl_grid(pk, dev, ("y1", bi), rows, cpad_o, torch.bfloat16)
        _lib.call("gldm_conv3d_tc_cl", x_cl.data_ptr(), w1_img.data_ptr(), blk["b1"].data_ptr(), B, ci, co, r, y1.data_ptr(),
                  0, cpad_o, stats[0].data_ptr(), ws.data_ptr(), st)
        _lib.call("gldm_gn_swish_cl", y1.data_ptr(), 0, cpad_o, stats[0].data_ptr(), blk["g1w"].data_ptr(),
                  blk["g1b"].data_ptr(), B, co, r, blk["eps1"], None, None, st)
    elif co == 48:       # 3-channel input: strict-fp32 SIMT Conv3d straight into the channels-last form + statistics
        y1 = _cl_grid(pk, dev, ("y1", bi), rows, cpad_o, torch.bfloat16)
        _lib.call("gldm_conv3d_k3_f32_cl", grid.data_ptr(), blk["w1"].data_ptr(), blk["b1"].data_ptr(), B, ci, r,
                  y1.data_ptr(), cpad_o, stats[0].data_ptr(), ws.data_ptr(), st)
        _lib.call("gldm_gn_swish_cl", y1.data_ptr(), 0, cpad_o, stats[0].data_ptr(), blk["g1w"].data_ptr(),
                  blk["g1b"].data_ptr(), B, co, r, blk["eps1"], None, None, st)
    else:
        t = torch.empty((B, co, r3), device=dev, dtype=torch.float32)
        _lib.call("gldm_conv3d_k3_f32", grid.data_ptr(), blk["w1"].data_ptr(), blk["b1"].data_ptr(), B, ci, co, r,
                  t.data_ptr(), st)
        _lib.call("gldm_groupnorm_swish_f32", t.data_ptr(), blk["g1w"].data_ptr(), blk["g1b"].data_ptr(), B, co, r3,
                  blk["groups"], blk["eps1"], None, st)
        y1 = _cl_grid(pk, dev, ("y1", bi), rows, cpad_o, torch.bfloat16)
        _lib.call("gldm_cl_pad", t.data_ptr(), B, co, r, y1.data_ptr(), st)
    # grid the devoxelisation reads: bf16 rows like every other activation of this path (half the traffic of the second
    # GroupNorm pass and of the corner gathers; encoder latent 3.5e-4 from the fp32 path instead of 3.1e-4), or fp32 rows
    # with GLDM_DEVOX_BF16=0
    y2_f32 = 1 if os.environ.get("GLDM_DEVOX_BF16") == "0" else 0
    y2_stride = co if y2_f32 else cpad_o
    y2 = _cl_grid(pk, dev, ("y2", bi, y2_f32), rows, y2_stride, torch.float32 if y2_f32 else torch.bfloat16)
    _lib.call("gldm_conv3d_tc_cl", y1.data_ptr(), w2_img.data_ptr(), blk["b2"].data_ptr(), B, co, co, r, y2.data_ptr(), y2_f32,
              y2_stride, stats[1].data_ptr(), ws.data_ptr(), st)
    _lib.call("gldm_gn_swish_cl", y2.data_ptr(), y2_f32, y2_stride, stats[1].data_ptr(), blk["g2w"].data_ptr(),
              blk["g2b"].data_ptr(), B, co, r, blk["eps2"], se_sum.data_ptr(), ws.data_ptr(), st)
    gate = torch.empty((B, co), device=dev, dtype=torch.float32)
    _lib.call("gldm_se_gate_sum", se_sum.data_ptr(), r3, blk["se1"].data_ptr(), blk["se2"].data_ptr(), B, co,
              blk["se1"].shape[0], gate.data_ptr(), st)
    pt = _pw(feats, blk["pw"], blk["pscale"], blk["pshift"], None, 1)
    fused = torch.empty((B, co, N), device=dev, dtype=torch.float32)
    _lib.call("gldm_devox_cl", norm.data_ptr(), y2.data_ptr(), y2_f32, y2_stride, gate.data_ptr(), pt.data_ptr(), B, co, N, r,
              fused.data_ptr(), st)
    return fused


def _encoder_pass(pk, xyz, precision="fp32"):
    dev = xyz.device
    st = _stream(dev)
    with torch.cuda.device(dev):
        B, N, _ = xyz.shape
        feats = xyz.float().transpose(1, 2).contiguous()       # [B,3,N] (layout only)
        coords = feats
        for bi, blk in enumerate(pk.blocks):
            if blk["kind"] == "pvconv" and precision == "bf16" and pk.tc_conv_weights()[bi][1] is not None \
                    and blk["groups"] == 8 and blk["cout"] % 16 == 0:
                feats = _pvconv_voxel_branch_tc(pk, bi, blk, feats, coords, B, N, st, dev)
            elif blk["kind"] == "pvconv":
                r, ci, co = blk["r"], blk["cin"], blk["cout"]
                tcw = pk.tc_conv_weights()[bi] if precision == "bf16" else (None, None)
                r3 = r ** 3
                grid = torch.empty((B, ci, r3), device=dev, dtype=torch.float32)
                norm = torch.empty((B, 3, N), device=dev, dtype=torch.float32)
                _lib.call("gldm_voxelize_fused", feats.data_ptr(), coords.data_ptr(), B, ci, N, r, grid.data_ptr(),
                          norm.data_ptr(), None, st)
                y1 = torch.empty((B, co, r3), device=dev, dtype=torch.float32)
                w1_img = None if getattr(tcw[0], "narrow_k", False) else tcw[0]      # (the K = 16 image is for the fused path)
                _conv3d(grid, blk["w1"], w1_img, blk["b1"], B, ci, co, r, y1, st)
                _lib.call("gldm_groupnorm_swish_f32", y1.data_ptr(), blk["g1w"].data_ptr(), blk["g1b"].data_ptr(), B, co,
                          r3, blk["groups"], blk["eps1"], None, st)
                y2 = torch.empty((B, co, r3), device=dev, dtype=torch.float32)
                _conv3d(y1, blk["w2"], tcw[1], blk["b2"], B, co, co, r, y2, st)
                se_mean = torch.empty((B, co), device=dev, dtype=torch.float32)
                _lib.call("gldm_groupnorm_swish_f32", y2.data_ptr(), blk["g2w"].data_ptr(), blk["g2b"].data_ptr(), B, co,
                          r3, blk["groups"], blk["eps2"], se_mean.data_ptr(), st)
                gate = torch.empty((B, co), device=dev, dtype=torch.float32)
                _lib.call("gldm_se_gate_f32", se_mean.data_ptr(), blk["se1"].data_ptr(), blk["se2"].data_ptr(), B, co,
                          blk["se1"].shape[0], gate.data_ptr(), st)
                pt = _pw(feats, blk["pw"], blk["pscale"], blk["pshift"], None, 1)
                fused = torch.empty((B, co, N), device=dev, dtype=torch.float32)
                _lib.call("gldm_devox_gate_add_f32", norm.data_ptr(), y2.data_ptr(), gate.data_ptr(), pt.data_ptr(), B,
                          co, N, r, fused.data_ptr(), st)
                feats = fused
            elif precision == "fp32":
                feats = _pw(feats, blk["pw"], blk["pscale"], blk["pshift"], None, 1)
        if precision == "fp32":
            h = _pw(feats, pk.wd, None, pk.bd, None, 0)
            h = _pw(h, pk.wo, None, pk.bo, None, 0)             # [B, C_out, N]
        else:
            h = _pointwise_tail_tc(pk, feats, st)               # tensor-core chain, [B, C_out, N] fp32
        z = torch.empty((B, pk.out_channels, pk.out_features), device=dev, dtype=torch.float32)
        _lib.call("gldm_linear_lastdim_f32", h.data_ptr(), pk.wl.data_ptr(), pk.bl.data_ptr(), B * pk.out_channels, N,
                  pk.out_features, z.data_ptr(), st)
    return z
