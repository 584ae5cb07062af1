import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np, torch
from oracle import torch_ops as T
import test_gpu_fused as tg
cuda = torch.device('cuda:0')
def rel(a, b): return ((a.double() - b.double()).norm() / max(b.double().norm().item(), 1e-30)).item()
for variance in (0.3,):
    ds, osdf, odev, orend, tr, batch_cpu = tg._setup(cuda, variance=variance)
    o, d, pn, vinv, nrm, msk = batch_cpu
    batch, near, far = tg._to_gpu_batch(ds, batch_cpu, cuda)
    step = 0.02
    jitter = torch.rand(o.shape[0])
    orend.sampling_step_size = step
    out = orend.render(o, d, pn, near, far, vinv, jitter=jitter)
    loss, parts = T.losses(out, nrm, msk)
    loss.backward()
    tr.forward_backward(batch, step, jitter.to(cuda))
    S = out["n_samples"]; dm = out["diff_mask"]; E = int(dm.sum())
    b = tr.buf
    print("S", S, "E", E, "fused", b.totals.tolist())
    sdf_o = out["sdf_all"].detach().reshape(-1, 9)
    sdf_f = b.sdf[:9 * (S + E)].view(-1, 9).cpu()
    print("sdf max abs err", (sdf_f - sdf_o).abs().max().item(), "rel", rel(sdf_f, sdf_o))
    # forward pieces via API outputs
    import ctypes as C
    from supernormal_b200._lib import call, ptr
    from supernormal_b200.trainer import make_batch_struct
    bs = make_batch_struct(batch["rays_o"], batch["rays_d"], batch["plane_n"], batch["near"], batch["far"], batch["v_inv"], batch["normal_gt"], batch["mask"])
    net = tr.model.net_struct()
    gradients = torch.zeros(S, 9, 3, device=cuda); weights = torch.zeros(S, 9, device=cuda); st = torch.zeros(8, device=cuda)
    call("snb_render_fwd", C.byref(bs), C.byref(net), C.byref(b.struct), ptr(b.sdf), ptr(b.comp), ptr(b.wsum), ptr(gradients), ptr(weights), ptr(st))
    print("weights rel", rel(weights.cpu(), out["weights"].detach().reshape(S, 9)), "gradients rel", rel(gradients.cpu(), out["gradients"].detach().reshape(S, 9, 3)),
          "max", (gradients.cpu() - out["gradients"].detach().reshape(S, 9, 3)).abs().max().item())
    print("comp rel", rel(b.comp.cpu().view(-1, 3, 3, 3), out["comp_normal"].detach()), "wsum rel", rel(b.wsum.cpu().view(-1, 3, 3, 1), out["weight_sum"].detach()))
    # seeds
    g_o = out["sdf_all"].grad.reshape(-1, 9)
    d0, d1 = b.d_sdf0[:9 * S].view(S, 9).cpu(), b.d_sdf1[:9 * S].view(S, 9).cpu()
    es = b.end_slot[:S].cpu()
    g_f_start = d0.clone()
    link = (es[:-1] < 0)
    g_f_start[1:][link] += d1[:-1][link]
    g_f_end = d1[es >= 0]
    print("diff_mask == (end_slot>=0):", torch.equal(dm, es >= 0))
    print("d_sdf start rel", rel(g_f_start, g_o[:S]), "max", (g_f_start - g_o[:S]).abs().max().item(), g_o[:S].abs().max().item())
    print("d_sdf end rel", rel(g_f_end, g_o[S:]), "max", (g_f_end - g_o[S:]).abs().max().item(), g_o[S:].abs().max().item())
    # where are the start errors
    err = (g_f_start - g_o[:S]).abs()
    idx = err.flatten().topk(8).indices
    pidx = out["samples"][0]
    for i in idx.tolist():
        s_, k_ = divmod(i, 9)
        print("  s", s_, "k", k_, "patch", int(pidx[s_]), "err", err[s_, k_].item(), "ref", g_o[s_, k_].item(), "got", g_f_start[s_, k_].item(), "alpha", out["alpha"].reshape(S, 9)[s_, k_].item(),
              "w", out["weights"].reshape(S,9)[s_,k_].item(), "first/last in patch", bool(s_ == 0 or pidx[s_-1] != pidx[s_]), bool(s_ == S-1 or pidx[s_+1] != pidx[s_]))
