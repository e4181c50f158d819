"""Generate tests/golden/*.npz from the UNMODIFIED reference (/root/reference) and pin the oracle to it.

Run in the build container only (the reference does not exist on the GPU box):
    python oracle/make_golden.py

For every stage of the hot path this imports the reference function, runs it on seeded
synthetic inputs (dm-nerf_b200/synth.py), asserts that oracle/dmnerf_oracle.py reproduces the
reference output bit-for-bit on this machine, and stores inputs + REFERENCE outputs as small fixtures.
Weights are not stored: they are regenerated from the recorded PCG64 seed by synth.make_weights.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, "/root/reference")

from networks.render import dm_nerf as ref_dm_nerf, render_train as ref_render_train   # noqa: E402
from networks.dm_nerf import get_embedder as ref_get_embedder, DM_NeRF as RefNet       # noqa: E402
from networks.helpers import sample_pdf as ref_sample_pdf, z_val_sample as ref_z_val_sample, \
    get_rays_k as ref_get_rays_k                                                            # noqa: E402

torch.autograd.set_detect_anomaly(False)   # reference networks/dm_nerf.py:5 turns it on at import

from oracle import dmnerf_oracle as O   # noqa: E402
from dmnerf_b200 import synth          # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def ref_net(weights_np, ins_num):
    net = RefNet(8, 256, 63, 27, [4], ins_num)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in weights_np.items()})
    return net


def same(a, b, what):
    a, b = a.detach(), b.detach()
    if not torch.equal(a, b):
        d = (a - b).abs().max().item()
        raise SystemExit("oracle != reference for %s (max abs diff %g)" % (what, d))


def t2n(x):
    return x.detach().cpu().numpy()


def main():
    g = np.random.Generator(np.random.PCG64(1234))

    # ---- a4: Embedder.embed ------------------------------------------------------------------
    x = (g.standard_normal((257, 3)) * 6.0).astype(np.float32)
    x[0] = [20.0, -19.5, 0.0]
    xt = torch.from_numpy(x)
    pe, pdim = ref_get_embedder(10)
    ve, vdim = ref_get_embedder(4)
    assert (pdim, vdim) == (63, 27)
    e_pos, e_dir = pe.embed(xt), ve.embed(xt / xt.norm(dim=-1, keepdim=True))
    same(O.embed(xt, 10), e_pos, "embed L=10")
    same(O.embed(xt / xt.norm(dim=-1, keepdim=True), 4), e_dir, "embed L=4")
    np.savez(os.path.join(OUT, "embed.npz"), x=x, pos=t2n(e_pos), dir=t2n(e_dir))

    # ---- a5: DM_NeRF.forward -------------------------------------------------------------------
    for ins_num, seed, M in ((13, 11, 300), (59, 12, 130)):
        w = synth.make_weights(seed, ins_num)
        net = ref_net(w, ins_num)
        pts = (g.standard_normal((M, 3)) * 3.0).astype(np.float32)
        dirs = g.standard_normal((M, 3)).astype(np.float32)
        dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
        xin = torch.cat([pe.embed(torch.from_numpy(pts)), ve.embed(torch.from_numpy(dirs))], -1)
        with torch.no_grad():
            y = net(xin)
        same(O.mlp_forward(O.to_torch(w), xin), y, "mlp ins_num=%d" % ins_num)
        np.savez(os.path.join(OUT, "mlp_ins%d.npz" % ins_num), seed=seed, ins_num=ins_num,
                 x=t2n(xin), y=t2n(y))

    # ---- a6: render_train ----------------------------------------------------------------------
    N, S, C = 24, 192, 18
    raw = (g.standard_normal((N, S, C)) * 2.0).astype(np.float32)
    raw[:, :, 3] = (g.standard_normal((N, S)) * 3.0 - 0.5).astype(np.float32)
    raw[0, :, 3] = -1.0                      # fully empty ray: all weight mass stays zero
    raw[1, :, 3] = 50.0                      # opaque at the first sample
    z = np.sort(g.uniform(4.0, 15.0, size=(N, S)).astype(np.float32), -1)
    z[2, 10:14] = z[2, 10]                   # repeated depths (zero-length intervals)
    rd = (g.standard_normal((N, 3)) * 1.3).astype(np.float32)
    rgb, wts, dep, ins = ref_render_train(torch.from_numpy(raw), torch.from_numpy(z), torch.from_numpy(rd))
    o = O.composite(torch.from_numpy(raw), torch.from_numpy(z), torch.from_numpy(rd))
    for a, b, n in zip(o[:4], (rgb, wts, dep, ins), ("rgb", "weights", "depth", "ins")):
        same(a, b, "composite " + n)
    np.savez(os.path.join(OUT, "composite.npz"), raw=raw, z=z, rays_d=rd, rgb=t2n(rgb), weights=t2n(wts),
             depth=t2n(dep), ins=t2n(ins))

    # ---- a7: sample_pdf (det and random u) -------------------------------------------------------
    N = 40
    zc = t2n(ref_z_val_sample(N, 4.0, 15.0, 64)).copy()
    bins = 0.5 * (zc[:, 1:] + zc[:, :-1])
    wts = g.uniform(0, 1, size=(N, 62)).astype(np.float32) ** 8      # peaky
    wts[0] = 0.0                                                       # all-zero weights -> uniform pdf
    wts[1, :] = 0.0; wts[1, 30] = 1.0                                  # single spike (denom < 1e-5 elsewhere)
    wts[2, :] = 1e-7                                                   # tiny weights
    det = ref_sample_pdf(torch.from_numpy(bins), torch.from_numpy(wts), 128, det=True)
    same(O.sample_pdf(torch.from_numpy(bins), torch.from_numpy(wts), 128, det=True), det, "sample_pdf det")
    torch.manual_seed(77)
    u = torch.rand(N, 128)
    torch.manual_seed(77)
    rnd = ref_sample_pdf(torch.from_numpy(bins), torch.from_numpy(wts), 128, det=False)
    same(O.sample_pdf(torch.from_numpy(bins), torch.from_numpy(wts), 128, det=False, u=u), rnd, "sample_pdf rand")
    np.savez(os.path.join(OUT, "sample_pdf.npz"), bins=bins, weights=wts, det=t2n(det), u=t2n(u), rnd=t2n(rnd))

    # ---- a9 + f1: z_val_sample, get_rays_k --------------------------------------------------------
    wl = synth.workload("dmsr_study")
    ro, rd_ = ref_get_rays_k(wl["H"], wl["W"], torch.from_numpy(wl["K"]), torch.from_numpy(wl["c2w"]))
    oo, od = O.get_rays_k(wl["H"], wl["W"], torch.from_numpy(wl["K"]), torch.from_numpy(wl["c2w"]))
    same(od, rd_, "get_rays_k d"); same(oo, ro, "get_rays_k o")
    same(O.z_val_sample(5, 4.0, 15.0, 64), ref_z_val_sample(5, 4.0, 15.0, 64), "z_val_sample")
    idx = np.array([0, 1, 639, 640, 153600, 307199])
    np.savez(os.path.join(OUT, "rays.npz"), K=wl["K"], c2w=wl["c2w"], idx=idx,
             rays_o=t2n(ro.reshape(-1, 3))[idx], rays_d=t2n(rd_.reshape(-1, 3))[idx],
             z=t2n(ref_z_val_sample(1, 4.0, 15.0, 64))[0],
             z_replica=t2n(ref_z_val_sample(1, 0.0, 6.5, 64))[0])

    # ---- a1: dm_nerf end to end (inference + train-mode perturb), + C4 gradients -------------------
    for tag, wlname, ins_num, N in (("study", "dmsr_study", 13, 16), ("room0", "replica_room0", 59, 6)):
        wl = synth.workload(wlname)
        sel = np.linspace(0, wl["H"] * wl["W"] - 1, N).astype(np.int64)
        ro_, rd_ = torch.from_numpy(wl["rays_o"][sel]), torch.from_numpy(wl["rays_d"][sel])
        wc, wf = synth.make_weights(101, ins_num), synth.make_weights(202, ins_num)
        nc, nf = ref_net(wc, ins_num), ref_net(wf, ins_num)
        zc = ref_z_val_sample(N, wl["near"], wl["far"], 64)
        args = types.SimpleNamespace(perturb=0.0, N_importance=128, is_train=False, N_ins=None)
        with torch.no_grad():
            ref = ref_dm_nerf(torch.stack([ro_, rd_], 0), pe, ve, nc, nf, zc, args)
            mine = O.render(ro_, rd_, O.to_torch(wc), O.to_torch(wf), zc, perturb=0.0)
        for k in ref:
            same(mine[k], ref[k], "dm_nerf[%s] %s" % (tag, k))
        save = dict(sel=sel, rays_o=t2n(ro_), rays_d=t2n(rd_), near=wl["near"], far=wl["far"], ins_num=ins_num,
                    seed_coarse=101, seed_fine=202)
        for k, v in ref.items():
            save["det_" + k] = t2n(v)

        # train mode: perturb=1 with the reference's own torch.rand draws, gradients of the C4 loss
        args = types.SimpleNamespace(perturb=1.0, N_importance=128, is_train=True, N_ins=None)
        torch.manual_seed(5)
        t_rand = torch.rand(N, 64)
        u = torch.rand(N, 128)
        torch.manual_seed(5)
        target = torch.from_numpy(g.uniform(0, 1, size=(N, 3)).astype(np.float32))
        ref = ref_dm_nerf(torch.stack([ro_, rd_], 0), pe, ve, nc, nf, zc, args)
        loss = O.train_loss(ref, target)
        loss.backward()
        pc, pf = O.to_torch(wc), O.to_torch(wf)
        for d in (pc, pf):
            for v in d.values():
                v.requires_grad_(True)
        mine = O.render(ro_, rd_, pc, pf, zc, perturb=1.0, t_rand=t_rand, u=u, is_train=True)
        for k in ref:
            same(mine[k], ref[k], "dm_nerf-train[%s] %s" % (tag, k))
        myloss = O.train_loss(mine, target)
        myloss.backward()
        save.update(t_rand=t2n(t_rand), u=t2n(u), target=t2n(target), loss=t2n(loss))
        for k, v in ref.items():
            save["trn_" + k] = t2n(v)
        for net, d, nm in ((nc, pc, "coarse"), (nf, pf, "fine")):
            for k, prm in net.named_parameters():
                gr = prm.grad if prm.grad is not None else torch.zeros_like(prm)
                mg = d[k].grad if d[k].grad is not None else torch.zeros_like(d[k])
                same(mg, gr, "grad %s %s" % (nm, k))
                # store full bias grads and a strided slice of each weight grad (keeps fixtures small)
                if k.endswith("bias"):
                    save["grad_%s_%s" % (nm, k)] = t2n(gr)
                else:
                    save["grad_%s_%s" % (nm, k)] = t2n(gr)[::8, ::8].copy()
                    save["gnorm_%s_%s" % (nm, k)] = np.float32(gr.norm().item())
        np.savez_compressed(os.path.join(OUT, "render_%s.npz" % tag), **save)

    # ---- the recipe quoted in SURVEY.md 8c (regenerated, not trusted) -------------------------------
    torch.manual_seed(0)
    nc, nf = RefNet(8, 256, 63, 27, [4], 13), RefNet(8, 256, 63, 27, [4], 13)
    gen = torch.Generator().manual_seed(1)
    d = torch.randn(1024, 3, generator=gen)
    d = 1.3 * d / d.norm(dim=-1, keepdim=True)
    o_ = torch.tensor([0.5, -2.0, 1.0]).expand(1024, 3)
    args = types.SimpleNamespace(perturb=0.0, N_importance=128, is_train=False, N_ins=None)
    with torch.no_grad():
        ref = ref_dm_nerf(torch.stack([o_, d], 0), pe, ve, nc, nf, ref_z_val_sample(1024, 4.0, 15.0, 64), args)
    print("survey recipe: rgb_fine[0] =", ref["rgb_fine"][0].tolist(), " z_fine[0,:3] =",
          ref["z_vals_fine"][0, :3].tolist(), " sum depth_fine =", ref["depth_fine"].double().sum().item())

    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden fixtures written to %s (%.1f KB total); oracle == reference bit-for-bit on all stages" % (OUT, tot / 1024))


if __name__ == "__main__":
    main()
