"""Development helper: run the row-major sampler kernel once (single denoiser evaluation) with progress markers and an
optional TMEM dump at a chosen job, and compare against torch."""
import os, sys, time
os.environ["GLDM_TC_ROWS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F
import _models
from graspldm_b200 import _lib
from oracle import model_torch as M
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dbg_job = int(sys.argv[2]) if len(sys.argv) > 2 else -1
m = _models.build("fpc").to(dev)
vae, ddm = _models.split_state_dicts(m)
prog = torch.zeros(16 + 128 * 512 // 2, dtype=torch.int64, device=dev)
prog[8] = dbg_job + 1
_lib.call("gldm_sampler_tc_set_profile", prog.data_ptr())
gen = torch.Generator().manual_seed(11)
x, zc = torch.randn(B, 1, 4, generator=gen), torch.randn(B, 3, 64, generator=gen)
tt = torch.randint(0, 1000, (B,), generator=gen)
f32 = m.diffusion_model.model(x.to(dev), time=tt.to(dev), z_cond=zc.to(dev))
got = m.diffusion_model.model(x.to(dev), time=tt.to(dev), z_cond=zc.to(dev), precision="bf16")
torch.cuda.synchronize()
print("max|bf16 rows - fp32|", (got - f32).abs().max().item(), "max|eps|", f32.abs().max().item())
if dbg_job >= 0:
    dump = prog[16:].view(torch.float32).view(128, 512).cpu()
    D = dump.view(4, 32, 512)[:, :B]            # [pos][sample][col]
    sd, p = ddm, "diffusion_model.model."
    with torch.no_grad():
        emb = M.time_embedding(sd, p, tt)
        inp = F.silu(F.linear(zc, sd[p + "input_emb_layers.0.weight"], sd[p + "input_emb_layers.0.bias"]))
        emb = emb.unsqueeze(-2).repeat(1, 3, 1) + inp
        x0 = F.conv1d(x, sd[p + "init_conv.weight"], sd[p + "init_conv.bias"], padding=3)     # [B,4,4]
        b = p + "blocks.0.0."
        def wsw(w):
            mean = w.mean(dim=(1, 2), keepdim=True); var = w.var(dim=(1, 2), unbiased=False, keepdim=True)
            return (w - mean) * torch.rsqrt(var + 1e-5)
        bf = lambda t: t.bfloat16().float()
        def show(name, got_acc, want):
            print(f"{name}: max err {(got_acc - want).abs().max().item():.4e}  max|want| {want.abs().max().item():.3f}")
        def acc(nch):
            return D[:, :, :nch].permute(1, 2, 0)                                             # [B, ch, pos]
        # stage-0 chain (4 channels)
        def rb_parts(bp, xin):
            e = F.linear(F.silu(emb), sd[bp + "mlp.1.weight"], sd[bp + "mlp.1.bias"]).transpose(1, 2)
            h1 = M._block(sd, bp + "block1.", xin, 4, e.chunk(2, dim=1))
            h2 = M._block(sd, bp + "block2.", h1, 4)
            return h1, h2 + xin
        h1, r0 = rb_parts(b, x0)
        b1 = p + "blocks.0.1."
        h1b, r1 = rb_parts(b1, r0)
        a = p + "blocks.0.2."
        xn = M._chan_layernorm(r1, sd[a + "fn.norm.g"])
        qkv = F.conv1d(bf(xn), bf(sd[a + "fn.fn.to_qkv.weight"]))
        att = M._linear_attention(sd, a, r1)
        if dbg_job == 1:
            show("s0.rb0.c2 acc", acc(4), F.conv1d(bf(h1), bf(wsw(sd[b + "block2.proj.weight"])), None, padding=1))
        if dbg_job == 2:
            show("s0.rb1.c1 acc", acc(4), F.conv1d(bf(r0), bf(wsw(sd[b1 + "block1.proj.weight"])), None, padding=1))
            print("res", D[:, 0, 384:388].t(), r0[0])
        if dbg_job == 3:
            show("s0.rb1.c2 acc", acc(4), F.conv1d(bf(h1b), bf(wsw(sd[b1 + "block2.proj.weight"])), None, padding=1))
        if dbg_job == 4:
            show("s0.qkv acc", acc(384), qkv)
            print("res", D[:, 0, 384:388].t(), r1[0])
        if dbg_job == 5:
            q, k, v = qkv.view(B, 3, 4, 32, 4).unbind(1)
            q = q.softmax(dim=-2) * 32 ** -0.5; k = k.softmax(dim=-1)
            ctx = torch.einsum("bhdn,bhen->bhde", k, v)
            out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(B, 128, 4)
            show("s0.out acc", acc(4), F.conv1d(bf(out), bf(sd[a + "fn.fn.to_out.0.weight"])))
        if dbg_job == 6:
            show("s0.down acc", acc(32), F.conv1d(bf(att), bf(sd[p + "blocks.0.3.weight"]), None, padding=1))
            print(acc(32)[0, :4], F.conv1d(bf(att), bf(sd[p + "blocks.0.3.weight"]), None, padding=1)[0, :4])
        if dbg_job == 0:
            w = sd[b + "block1.proj.weight"]
            mean = w.mean(dim=(1, 2), keepdim=True); var = w.var(dim=(1, 2), unbiased=False, keepdim=True)
            wn = (w - mean) * torch.rsqrt(var + 1e-5)
            y = F.conv1d(x0.bfloat16().float(), wn.bfloat16().float(), None, padding=1)      # without bias
            got_acc = D[:, :, :4].permute(1, 2, 0)                                            # [B, ch, pos]
            print("job0 acc max err", (got_acc - y).abs().max().item(), "max", y.abs().max().item())
            print(got_acc[0]); print(y[0])
            e = F.linear(F.silu(emb), sd[b + "mlp.1.weight"], torch.zeros_like(sd[b + "mlp.1.bias"])).sum(1)  # [B, 8]
            print("film scale", D[0, :, 128:132][:2], e[:2, :4]); print("film shift", D[0, :, 256:260][:2], e[:2, 4:])
            print("res", D[:, 0, 384:388].t(), x0[0])
        # stage outputs (= residual stream when the first job of the next stage is ready)
        xs, xcur = [], x0
        for i in range(4):
            bb = p + f"blocks.{i}."
            xcur = M._resnet_block(sd, bb + "0.", xcur, emb, 4)
            xcur = M._resnet_block(sd, bb + "1.", xcur, emb, 4)
            xcur = M._linear_attention(sd, bb + "2.", xcur)
            xcur = F.conv1d(xcur, sd[bb + "3.weight"], sd[bb + "3.bias"], padding=1)
            xs.append(xcur)
        for i, jb in enumerate((7, 14, 21)):
            if dbg_job == jb:
                nch = xs[i].shape[1]
                got_res = D[:, :, 384:384 + nch].permute(1, 2, 0)
                show(f"stage {i} output (residual at job {jb})", got_res, xs[i])
                bb = p + f"blocks.{i + 1}.0."
                show("  first conv acc", acc(nch), F.conv1d(bf(xs[i]), bf(wsw(sd[bb + "block1.proj.weight"])), None, padding=1))
        if dbg_job == 28:
            raw = D[:, :, 384:512].contiguous().view(torch.int32)
            lo = (raw << 16).view(torch.float32); hi = (raw & -65536).view(torch.float32)
            got_res = torch.stack((lo, hi), -1).reshape(4, B, 256).permute(1, 2, 0)
            show("stage 3 output (packed residual at job 28)", got_res, xs[3])
            bb = p + "final_res_block."
            show("  fin.c1 half 0 acc", acc(128), F.conv1d(bf(xs[3]), bf(wsw(sd[bb + "block1.proj.weight"])), None, padding=1)[:, :128])
        if dbg_job == 29:
            bb = p + "final_res_block."
            show("  fin.c1 half 1 acc", acc(128), F.conv1d(bf(xs[3]), bf(wsw(sd[bb + "block1.proj.weight"])), None, padding=1)[:, 128:])

