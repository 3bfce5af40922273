"""Golden-vector generator (runs ONLY in the build container, where /root/reference exists).

Imports the UNMODIFIED reference python modules from /root/reference, builds the models of the
named configs from seeded random init, runs the reference forward passes on CPU and stores small
input/output fixtures next to this file.  Three things stand in for packages/hardware that are
not available here (all under tests/golden/_shims or patched below, never shipped in the product):

  * addict.Dict, yapf.FormatCode           - tiny stand-ins (config plumbing only)
  * diffusers.DDPMScheduler/DDIMScheduler  - oracle/schedulers.py restatement (parity unpinned)
  * the CUDA-only `_pvcnn_backend` ops      - oracle/ops_np.py (the reference has no CPU path;
                                              oracle/ops_np.py itself is pinned on the GPU box
                                              against oracle/_ref, see make_golden_gpu.py)

Usage:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ops_np  # noqa: E402
import _models  # noqa: E402  (trained_like_: shared with the tests so both module trees get identical values)


def _install_cpu_backend():
    """Replace the reference's JIT-built CUDA module by a CPU object backed by oracle/ops_np.py."""
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))

    class _Backend:
        @staticmethod
        def avg_voxelize_forward(f, c, r):
            return tuple(t(a) for a in ops_np.avg_voxelize_forward(f.numpy(), c.numpy(), r))

        @staticmethod
        def trilinear_devoxelize_forward(r, training, c, f):
            return tuple(t(a) for a in ops_np.trilinear_devoxelize_forward(r, training, c.numpy(), f.numpy()))

        @staticmethod
        def furthest_point_sampling(c, m):
            return t(ops_np.furthest_point_sampling(c.numpy(), m))

        @staticmethod
        def gather_features_forward(f, i):
            return t(ops_np.gather_features_forward(f.numpy(), i.numpy()))

        @staticmethod
        def ball_query(c, p, r, u):
            return t(ops_np.ball_query(c.numpy(), p.numpy(), r, u))

        @staticmethod
        def grouping_forward(f, i):
            return t(ops_np.grouping_forward(f.numpy(), i.numpy()))

        @staticmethod
        def three_nearest_neighbors_interpolate_forward(p, c, f):
            return tuple(t(a) for a in ops_np.three_nearest_neighbors_interpolate_forward(p.numpy(), c.numpy(), f.numpy()))

    name = "grasp_ldm.models.modules.ext.pvcnn.modules.functional.backend"
    mod = types.ModuleType(name)
    mod._backend = _Backend()
    mod.__all__ = ["_backend"]
    sys.modules[name] = mod


_install_cpu_backend()

from grasp_ldm.models import GraspCVAE, GraspLatentDDM  # noqa: E402
from grasp_ldm.models.modules.resnets import TimeConditionedResNet1D  # noqa: E402
from grasp_ldm.utils.config import Config  # noqa: E402
from grasp_ldm.utils.rotations import tmrp_to_H  # noqa: E402

CONFIGS = {
    "fpc": f"{REF}/configs/generation/fpc/fpc_1a_latentc3_z4_pc64_180k.py",
    "ppc": f"{REF}/configs/generation/partial_pc/ppc_1a_partial_63cat8k_filtered_latentc3_z16_pc256_180k.py",
}


def build_reference_ldm(name, seed=0, scheduler="ddpm"):
    """Construction order of tools/inference.py:514-516: DDM first, then the VAE."""
    cfg = Config.fromfile(CONFIGS[name])
    cfg.model.ddm.model.args.noise_scheduler_type = scheduler
    # models/builder.py:59-91 builds nested `model=` dicts depth-first (denoiser, then the DDM);
    # builder.py itself cannot be imported here (it pulls in trimesh via grasp_classifier.py).
    torch.manual_seed(seed)
    ddm_args = dict(cfg.model.ddm.model.args)
    assert ddm_args["model"]["type"] == "TimeConditionedResNet1D"
    ddm_args["model"] = TimeConditionedResNet1D(**ddm_args["model"]["args"])
    model = GraspLatentDDM(**ddm_args)
    model.set_vae_model(GraspCVAE(**cfg.model.vae.model.args))
    return model.eval()


def synthetic_clouds(B, N=1024, seed=1234, dist="S"):
    """(S) sphere surface with per-axis scale, (G) randn*0.7."""
    g = torch.Generator().manual_seed(seed)
    if dist == "S":
        p = torch.randn(B, N, 3, generator=g)
        p = p / p.norm(dim=-1, keepdim=True)
        p = p * (0.5 + torch.rand(B, 1, 3, generator=g))
        return p - p.mean(1, keepdim=True)
    return torch.randn(B, N, 3, generator=g) * 0.7


def reference_functions(path, names):
    """Compile selected top-level functions / methods of a reference file without importing the module (inference_base.py
    imports models/builder.py -> trimesh, absent here).  The code that runs is the reference's own, read from R at
    generation time; nothing is copied into this repository."""
    import ast
    tree = ast.parse(open(path).read())
    out = {}
    ns = {"torch": torch, "Tensor": torch.Tensor, "np": np}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in out:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), ns)
            out[node.name] = ns[node.name]
    return out


def normalize_golden():
    """normalize_input -> (fixed model outputs) -> unnormalize_grasps -> tmrp_to_H on raw, un-centred clouds:
    inference_base.py:61-84, 103-130, 182-212."""
    fn = reference_functions(f"{REF}/grasp_ldm/inference/inference_base.py",
                             ("set_normalization_params", "normalize_input", "unnormalize_grasps", "unnormalize_pc"))
    norm = types.SimpleNamespace(pc_shift=[0.01, -0.02, 0.005], grasp_shift=[0.01, -0.02, 0.005, 0.1, -0.05, 0.2],
                                 translation_scale=0.05, rotation_scale=0.5)
    g = torch.Generator().manual_seed(21)
    raw = synthetic_clouds(3, seed=77, dist="S") * 0.05 + torch.tensor([[[0.4, -0.2, 0.9]], [[-1.5, 0.3, 0.05]], [[0.0, 2.0, -0.7]]])
    tmrp = torch.randn(3, 4, 6, generator=g)
    out = {}
    # single cloud [N,3]: inference_base.py (its in-place `grasp_mean[..., :3] += pc_mean` only broadcasts for one cloud)
    self = types.SimpleNamespace(device="cpu")
    fn["set_normalization_params"](self, norm)
    scale6 = self._INPUT_GRASP_SCALE.clone()
    pcn, metas = fn["normalize_input"](self, raw[1].clone())
    res = {"single": (pcn, metas, tmrp[1])}
    # batch [B,N,3]: tools/inference.py:570-591 (InferenceLDM.normalize_input), whose PC_MEAN / PC_STD / GRASP_MEAN /
    # GRASP_STD attributes are the same four constants
    fb = reference_functions(f"{REF}/tools/inference.py", ("normalize_input",))
    selfb = types.SimpleNamespace(PC_MEAN=torch.tensor(norm.pc_shift), PC_STD=torch.ones(3) * norm.translation_scale,
                                  GRASP_MEAN=torch.tensor(norm.grasp_shift), GRASP_STD=scale6)
    pcb, metab = fb["normalize_input"](selfb, raw.clone())
    res["batch"] = (pcb, metab, tmrp)
    for tag, (pcn, metas, t) in res.items():
        out[f"{tag}_pc"] = pcn.numpy()
        for k in ("pc_mean", "pc_std", "grasp_mean", "grasp_std"):
            out[f"{tag}_{k}"] = metas[k].numpy()
        tm = {k: v for k, v in metas.items() if isinstance(v, torch.Tensor)}
        if tag == "single":
            tm = {k: v.unsqueeze(0) for k, v in tm.items()}
            gu = fn["unnormalize_grasps"](t.unsqueeze(0), tm)[0]
        else:
            gu = fn["unnormalize_grasps"](t, tm)
        out[f"{tag}_grasp_tmrp"] = gu.numpy()
        out[f"{tag}_H"] = tmrp_to_H(gu).numpy()
        out[f"{tag}_pc_unnorm"] = fn["unnormalize_pc"](pcn, metas).numpy()
    np.savez_compressed(f"{HERE}/normalize_input.npz", raw=raw.numpy(), tmrp=tmrp.numpy(),
                        pc_shift=np.array(norm.pc_shift, np.float32), grasp_shift=np.array(norm.grasp_shift, np.float32),
                        translation_scale=np.float32(norm.translation_scale), rotation_scale=np.float32(norm.rotation_scale),
                        **out)


def edm_golden():
    """ElucidatedDiffusion.sample_normal / sample_using_dpmpp of the reference (elucidated_diffusion.py:179-315) around the
    reference denoiser with the seeded fpc weights; the N(0,1) draws are recorded by replaying the global CPU generator."""
    from grasp_ldm.models.diffusion.elucidated_diffusion import ElucidatedDiffusion
    cfg = Config.fromfile(CONFIGS["fpc"])
    torch.manual_seed(0)
    den = TimeConditionedResNet1D(**dict(cfg.model.ddm.model.args)["model"]["args"]).eval()     # same weights as build_reference_ldm
    edm = ElucidatedDiffusion(net=den, seq_length=4).eval()
    g = torch.Generator().manual_seed(33)
    B, n_heun, n_dpm = 6, 8, 10
    zc = torch.randn(B, 3, 64, generator=g)
    out = dict(z_cond=zc.numpy())
    with torch.no_grad():
        torch.manual_seed(101)
        x_init = torch.randn(B, 1, 4)
        noise = torch.stack([torch.randn(B, 1, 4) for _ in range(n_heun)])
        torch.manual_seed(101)
        x, _ = edm.sample(use_dpmpp=False, batch_size=B, z_cond=zc, num_sample_steps=n_heun)
        out.update(heun_x_init=x_init.numpy(), heun_noise=noise.numpy(), heun_x=x.numpy(), heun_steps=np.int64(n_heun))
        torch.manual_seed(202)
        x_init = torch.randn(B, 1, 4)
        torch.manual_seed(202)
        x, _ = edm.sample(use_dpmpp=True, batch_size=B, z_cond=zc, num_sample_steps=n_dpm)
        out.update(dpmpp_x_init=x_init.numpy(), dpmpp_x=x.numpy(), dpmpp_steps=np.int64(n_dpm))
        # one preconditioned evaluation at a few noise levels
        xs = torch.randn(B, 1, 4, generator=g)
        for k, sg in enumerate((80.0, 2.5, 0.05)):
            out[f"denoise_{k}"] = edm.preconditioned_network_forward(xs * sg, sg, z_cond=zc).numpy()
        out["denoise_x"] = xs.numpy()
    np.savez_compressed(f"{HERE}/edm_fpc.npz", **out)


def ppc_ldm_golden():
    """Partial-point-cloud model family (latent 16, conditioning width 256): 10 DDPM steps end to end through the
    unmodified reference classes, injected noise (same recipe as the fpc LDM fixtures in main())."""
    nobj, G, nsteps = 2, 3, 10
    m = build_reference_ldm("ppc", scheduler="ddpm")
    D = m.diffusion_model.n_dims
    m.set_inference_timesteps(nsteps)
    xyz = torch.cat([synthetic_clouds(2, seed=1234, dist="S"), synthetic_clouds(1, seed=99, dist="G")])
    gg = torch.Generator().manual_seed(42)
    noise = torch.randn(nsteps, nobj * G, 1, D, generator=gg)
    m.diffusion_model.noise_scheduler.injected_noise = list(noise)
    torch.manual_seed(42)
    x_T = torch.randn((nobj * G, 1, D))
    torch.manual_seed(42)
    with torch.no_grad():
        (tm, lg), _ = m.generate_grasps(xyz[:nobj], num_grasps=G, device="cpu")
    np.savez_compressed(f"{HERE}/ldm_ppc_ddpm10.npz", x_T=x_T.numpy(), noise=noise.numpy(), tmrp=tm.numpy(), logit=lg.numpy())


def trained_golden(man):
    """Checkpoint-like state (tests/_models.py::trained_like_: non-trivial BatchNorm running statistics and affine
    parameters of every BatchNorm / GroupNorm / LayerNorm) through the unmodified reference classes: single evaluations,
    encoder, VAE mode, LDM (100 DDPM / 10 DDIM steps) for both model families, and one BASELINE config-2-sized run
    (64 objects x 20 grasps, 100 DDPM steps) for fpc.  x_T / noise of the config-2 run are not stored: the test re-draws
    them from the same seeded CPU generators."""
    np_ = lambda x: x.detach().cpu().numpy()
    base = json.load(open(f"{HERE}/state_dict_manifest.json")) if not man else man
    for name in ("fpc", "ppc"):
        model = _models.trained_like_(build_reference_ldm(name))
        full = manifest(model.state_dict())
        man[name + "_trained"] = {k: v for k, v in full.items() if base[name][k]["sha"] != v["sha"]}
        D = model.diffusion_model.n_dims
        Dc = 64 if name == "fpc" else 256
        den, vae = model.diffusion_model.model, model.vae_model
        d0 = np.load(f"{HERE}/dense_{name}.npz")
        t = lambda k: torch.from_numpy(d0[k])
        xyz = torch.cat([synthetic_clouds(2, seed=1234, dist="S"), synthetic_clouds(1, seed=99, dist="G")])
        nobj, G = 2, 3
        with torch.no_grad():
            eps = den(t("x"), time=t("t"), z_cond=t("z_cond"))
            tmrp, logit = vae.decoder(t("z_h"), t("z_cond"))
            np.savez_compressed(f"{HERE}/dense_{name}_trained.npz", eps=np_(eps), tmrp=np_(tmrp), logit=np_(logit))
            np.savez_compressed(f"{HERE}/encoder_{name}_trained.npz", z_pc=np_(vae.encode_pc(xyz)))
            torch.manual_seed(5)
            z_h = torch.randn(nobj * G, D)
            torch.manual_seed(5)
            tm, lg = vae.generate_grasps(xyz[:nobj], num_grasps=G)
            np.savez_compressed(f"{HERE}/vae_{name}_trained.npz", z_h=np_(z_h), tmrp=np_(tm), logit=np_(lg))
            runs = [("ddpm", 100, nobj, G, xyz[:nobj]), ("ddim", 10, nobj, G, xyz[:nobj])]
            if name == "fpc":
                runs.append(("ddpm", 100, 64, 20, synthetic_clouds(64, seed=1234, dist="S")))
            for sched, nsteps, no, ng, clouds in runs:
                m = _models.trained_like_(build_reference_ldm(name, scheduler=sched))
                m.set_inference_timesteps(nsteps)
                gg = torch.Generator().manual_seed(42)
                noise = torch.randn(nsteps, no * ng, 1, D, generator=gg)
                m.diffusion_model.noise_scheduler.injected_noise = list(noise)
                torch.manual_seed(42)
                x_T = torch.randn((no * ng, 1, D))
                torch.manual_seed(42)
                (tm, lg), _ = m.generate_grasps(clouds, num_grasps=ng, device="cpu")
                if no == 64:
                    np.savez_compressed(f"{HERE}/ldm_{name}_trained_config2.npz", tmrp=np_(tm), logit=np_(lg))
                else:
                    np.savez_compressed(f"{HERE}/ldm_{name}_trained_{sched}{nsteps}.npz", x_T=np_(x_T), noise=np_(noise),
                                        tmrp=np_(tm), logit=np_(lg))
    return man


def class_conditioned_golden():
    """ClassTimeConditionedResNet1D (class_conditioned_resnet.py:9-122) with the fpc denoiser arguments: single evaluations
    and a 10-step DDPM run of the reference GaussianDiffusion1D whose kwargs carry metas["mode_cls"] (gaussian_diffusion.py:271)."""
    from grasp_ldm.models.diffusion.gaussian_diffusion import GaussianDiffusion1D
    from grasp_ldm.models.modules.class_conditioned_resnet import ClassTimeConditionedResNet1D
    cfg = Config.fromfile(CONFIGS["fpc"])
    torch.manual_seed(0)
    den = _models.trained_like_(ClassTimeConditionedResNet1D(**dict(cfg.model.ddm.model.args)["model"]["args"]), 4).eval()
    g = torch.Generator().manual_seed(17)
    B = 6
    x, zc = torch.randn(B, 1, 4, generator=g), torch.randn(B, 3, 64, generator=g)
    t = torch.tensor([0, 3, 250, 500, 990, 999])
    cls = torch.tensor([0.0, 1.0, 1.0, 0.0, 2.0, -1.0]).view(B, 1)
    with torch.no_grad():
        eps = den(x, time=t, z_cond=zc, cls_cond=cls)
        ddm_args = dict(cfg.model.ddm.model.args)
        gd = GaussianDiffusion1D(model=den, n_dims=4, num_steps=1000, loss_type="l2", beta_schedule="linear", beta_start=5e-5,
                                 beta_end=1e-3, noise_scheduler_type="ddpm", variance_type="fixed_large").eval()
        gd.set_inference_timesteps(10)
        noise = torch.randn(10, B, 1, 4, generator=g)
        gd.noise_scheduler.injected_noise = list(noise)
        torch.manual_seed(8)
        x_T = torch.randn((B, 1, 4))
        torch.manual_seed(8)
        x0, _ = gd.sample(z_cond=zc, batch_size=B, device="cpu", metas=dict(mode_cls=cls.view(B)))
    np.savez_compressed(f"{HERE}/cls_fpc.npz", x=x.numpy(), t=t.numpy(), z_cond=zc.numpy(), cls=cls.numpy(), eps=eps.numpy(),
                        x_T=x_T.numpy(), noise=noise.numpy(), x0=x0.numpy(),
                        cls_w=den.cls_embed[0].weight.detach().numpy(), cls_b=den.cls_embed[0].bias.detach().numpy())


def pointnet_golden():
    """Set-abstraction family (SURVEY.md finding 1: the FPS / ball-query / grouping / 3-NN side of the operator extension is
    reached through PointNetSAModule / PointNetFPModule, i.e. PVCNN2 and PointNet2SSG) and the grasp classifier
    (grasp_classifier.py:13-143): the unmodified reference classes over the CPU backend (oracle/ops_np.py), checkpoint-like
    state (trained_like_), seeded construction so that the product's mirrors reproduce the weights bit for bit."""
    sys.modules.setdefault("trimesh", types.ModuleType("trimesh"))      # utils/gripper.py imports it at module level only
    from grasp_ldm.models.modules.ext.pvcnn.modules.pointnet import PointNetFPModule, PointNetSAModule
    from grasp_ldm.models.modules.ext.pvcnn.pointnet2 import PointNet2SSG
    from grasp_ldm.models.modules.ext.pvcnn.pvcnn_base import PVCNN2
    from grasp_ldm.models.grasp_classifier import PointsBasedGraspClassifier
    np_ = lambda x: x.detach().cpu().numpy()
    g = torch.Generator().manual_seed(31)
    out, man = {}, {}
    pcs = torch.cat([synthetic_clouds(1, seed=1234, dist="S"), synthetic_clouds(1, seed=99, dist="G") * 0.6]).transpose(1, 2).contiguous()
    extra = torch.randn(2, 3, 1024, generator=g) * 0.5
    with torch.no_grad():
        # multi-radius set abstraction and feature propagation on their own
        torch.manual_seed(3)
        sa = _models.trained_like_(PointNetSAModule(num_centers=128, radius=[0.2, 0.4], num_neighbors=[16, 48], in_channels=5,
                                                    out_channels=[(16, 32), (24, 40)]), 5).eval()
        f5 = torch.randn(2, 5, 1024, generator=g)
        sa_f, sa_c = sa((f5, pcs))
        torch.manual_seed(4)
        fp = _models.trained_like_(PointNetFPModule(in_channels=72 + 5, out_channels=(32, 16)), 6).eval()
        fp_f, _ = fp((pcs, sa_c, sa_f, f5))
        out.update(sa_in=np_(f5), sa_features=np_(sa_f), sa_centers=np_(sa_c), fp_features=np_(fp_f))
        man["sa"], man["fp"] = manifest(sa.state_dict()), manifest(fp.state_dict())
        torch.manual_seed(0)
        ssg = _models.trained_like_(PointNet2SSG(width_multiplier=0.5), 7).eval()
        out["ssg_in"] = np_(torch.cat([pcs, extra], 1))
        out["ssg_out"] = np_(ssg(torch.cat([pcs, extra], 1)))
        man["ssg"] = manifest(ssg.state_dict())
        torch.manual_seed(0)
        p2 = _models.trained_like_(PVCNN2(extra_feature_channels=0, width_multiplier=0.5, voxel_resolution_multiplier=0.5), 8).eval()
        out["pvcnn2_out"] = np_(p2(pcs))
        man["pvcnn2"] = manifest(p2.state_dict())
        torch.manual_seed(0)
        cls = PointsBasedGraspClassifier(
            num_pc_points=1024 + 64,
            points_backbone_config=dict(type="PVCNN", args=dict(in_channels=3, extra_feature_channels=1, scale_channels=0.25,
                                                                scale_voxel_resolution=0.5, num_blocks=(1, 1, 1, 1))),
            loss_config=types.SimpleNamespace(classification_loss=dict(type="BCEClassificationLoss", args={})))
        cls = _models.trained_like_(cls, 9).eval()
        grasp_pts = torch.randn(2, 64, 3, generator=g) * 0.3
        _, preds = cls(pcs.transpose(1, 2).contiguous(), grasp_pts, compute_loss=False)
        out.update(cls_grasp_points=np_(grasp_pts), cls_preds=np_(preds))
        man["classifier"] = manifest(cls.state_dict())
    np.savez_compressed(f"{HERE}/pointnet_family.npz", coords=np_(pcs), **out)
    with open(f"{HERE}/pointnet_manifest.json", "w") as f:
        json.dump(man, f, indent=0, sort_keys=True)


def manifest(sd):
    out = {}
    for k, v in sd.items():
        a = v.detach().cpu().contiguous().numpy()
        out[k] = dict(shape=list(a.shape), dtype=str(a.dtype), sha=hashlib.sha256(a.tobytes()).hexdigest()[:16])
    return out


def main():
    torch.set_num_threads(8)
    np_ = lambda x: x.detach().cpu().numpy()
    only = set(sys.argv[1:])          # e.g. `make_golden.py trained` regenerates one section
    if only:
        if "trained" in only:
            man = json.load(open(f"{HERE}/state_dict_manifest.json"))
            man = trained_golden(man)
            with open(f"{HERE}/state_dict_manifest.json", "w") as f:
                json.dump(man, f, indent=0, sort_keys=True)
        for tag, fn in (("normalize", normalize_golden), ("edm", edm_golden), ("ppc_ldm", ppc_ldm_golden),
                        ("pointnet", pointnet_golden), ("cls", class_conditioned_golden)):
            if tag in only:
                fn()
        return
    man = {}
    for name in ("fpc", "ppc"):
        model = build_reference_ldm(name)
        man[name] = manifest(model.state_dict())
        D = model.diffusion_model.n_dims
        Dc = 64 if name == "fpc" else 256
        den = model.diffusion_model.model
        vae = model.vae_model
        g = torch.Generator().manual_seed(7)
        with torch.no_grad():
            # ---- denoiser forward (resnets.py:558-616)
            B = 6
            x = torch.randn(B, 1, D, generator=g)
            t = torch.tensor([0, 1, 10, 500, 990, 999])
            zc = torch.randn(B, 3, Dc, generator=g)
            eps = den(x, time=t, z_cond=zc)
            # ---- decoder (grasp_vae.py:401-436)
            zh = torch.randn(B, D, generator=g)
            tmrp, logit = vae.decoder(zh, zc)
            np.savez_compressed(f"{HERE}/dense_{name}.npz", x=np_(x), t=np_(t), z_cond=np_(zc), eps=np_(eps),
                                z_h=np_(zh), tmrp=np_(tmrp), logit=np_(logit))
            # ---- encoder (pc_encoders.py:87-115), sphere + gaussian clouds
            xyz = torch.cat([synthetic_clouds(2, seed=1234, dist="S"), synthetic_clouds(1, seed=99, dist="G")])
            z_pc = vae.encode_pc(xyz)
            np.savez_compressed(f"{HERE}/encoder_{name}.npz", z_pc=np_(z_pc))
            if name != "fpc":
                continue
            # ---- LDM sampling: reference loop + scheduler restatement, injected noise
            nobj, G = 2, 3
            for sched, nsteps in (("ddpm", 10), ("ddpm", 100), ("ddim", 5), ("ddpm", None)):
                m = build_reference_ldm(name, scheduler=sched)
                if nsteps:
                    m.set_inference_timesteps(nsteps)
                n_exec = nsteps if nsteps else 1000
                gg = torch.Generator().manual_seed(42)
                noise = torch.randn(n_exec, nobj * G, 1, D, generator=gg)
                m.diffusion_model.noise_scheduler.injected_noise = list(noise)
                torch.manual_seed(42)   # x_T comes from the global CPU generator (gaussian_diffusion.py:253)
                x_T = torch.randn((nobj * G, 1, D))
                torch.manual_seed(42)
                (tm, lg), _ = m.generate_grasps(xyz[:nobj], num_grasps=G, device="cpu")
                tag = f"{sched}{nsteps if nsteps else 'full'}"
                np.savez_compressed(f"{HERE}/ldm_{name}_{tag}.npz", x_T=np_(x_T), noise=np_(noise),
                                    tmrp=np_(tm), logit=np_(lg))
            # ---- VAE mode (grasp_vae.py:226-255)
            torch.manual_seed(5)
            z_h = torch.randn(nobj * G, D)
            torch.manual_seed(5)
            tm, lg = vae.generate_grasps(xyz[:nobj], num_grasps=G)
            # ---- pose post-processing (tools/inference.py:627-656, rotations.py:298-302)
            metas = dict(grasp_std=torch.tensor([[.05, .05, .05, .5, .5, .5]]), grasp_mean=torch.zeros(1, 6))
            g_un = tm.view(nobj, G, 6) * metas["grasp_std"].unsqueeze(-2) + metas["grasp_mean"].unsqueeze(-2)
            H = tmrp_to_H(g_un)
            np.savez_compressed(f"{HERE}/vae_{name}.npz", z_h=np_(z_h), tmrp=np_(tm), logit=np_(lg),
                                grasp_tmrp=np_(g_un), H=np_(H))
    man = trained_golden(man)
    with open(f"{HERE}/state_dict_manifest.json", "w") as f:
        json.dump(man, f, indent=0, sort_keys=True)
    normalize_golden()
    edm_golden()
    ppc_ldm_golden()
    pointnet_golden()
    class_conditioned_golden()
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
