"""Generate tests/golden/manipulator.npz from the UNMODIFIED reference (networks/manipulator.py:18-205) and pin the oracle
to it.  The reference module imports image / metric libraries that are not installed here (lpips, cv2, imageio, skimage) and
two sibling modules for its evaluation loops; none of them is touched by the four functions on the edit path, so they are
replaced by empty stand-ins for the import only.         python oracle/make_golden_manipulator.py"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
for name in ("lpips", "cv2", "imageio", "skimage", "skimage.metrics", "tools", "tools.visualizer", "networks.evaluator"):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            for attr in ("metrics", "to8b", "ins_eval", "render_label2img", "render_gt_label2img"):
                setattr(m, attr, None)
            sys.modules[name] = m
import networks.manipulator as RM                                       # noqa: E402
from networks.dm_nerf import DM_NeRF as RefNet, get_embedder as ref_get_embedder   # noqa: E402
from oracle import dmnerf_oracle as O                                    # noqa: E402
from dmnerf_b200 import synth                                            # noqa: E402

torch.autograd.set_detect_anomaly(False)
OUT = os.path.join(ROOT, "tests", "golden")


def ref_net(weights_np, ins_num):
    net = RefNet(8, 256, 63, 27, [4], ins_num)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in weights_np.items()})
    return net


def main():
    ins_num, n, S, NI = 13, 40, 16, 32
    wc, wf = synth.make_weights(31, ins_num), synth.make_weights(32, ins_num)
    # instance heads with a wide spread, so that every label (and the "empty" class) wins somewhere
    g = np.random.Generator(np.random.PCG64(5))
    for w in (wc, wf):
        w["ins_linear.weight"] = (w["ins_linear.weight"] * 400).astype(np.float32)
        w["ins_linear.bias"] = (0.3 * g.standard_normal(ins_num + 1)).astype(np.float32)
    nc, nf = ref_net(wc, ins_num), ref_net(wf, ins_num)
    pc, pf = O.to_torch(wc), O.to_torch(wf)
    pe, _ = ref_get_embedder(10)
    ve, _ = ref_get_embedder(4)
    wl = synth.workload("dmsr_study")
    sel = np.linspace(0, 307199, n).astype(np.int64)
    ro, rd = torch.from_numpy(wl["rays_o"][sel]), torch.from_numpy(wl["rays_d"][sel])
    ori = torch.stack([ro, rd], 0)
    # two rigid "moves": the target rays are the original rays seen from the moved object's frame
    tars = []
    for ang, sh in ((0.3, (0.4, -0.2, 0.1)), (-0.2, (-0.3, 0.1, 0.25))):
        c, s_ = np.cos(ang), np.sin(ang)
        R = torch.tensor([[c, -s_, 0], [s_, c, 0], [0, 0, 1]], dtype=torch.float32)
        tars.append(torch.stack([ro @ R.T + torch.tensor(sh), rd @ R.T], 0))
    f_tar = torch.stack(tars, 0)                                          # [T, 2, N, 3]
    labels = [2, 7]
    args = types.SimpleNamespace(N_samples=S, N_importance=NI, near=float(wl["near"]), far=float(wl["far"]), target_labels=labels)
    gen = torch.Generator().manual_seed(11)
    us = [torch.rand(n, NI, generator=gen) for _ in range(2 + len(tars))]

    # the reference draws its u inside sample_pdf from the global generator: feed it the same numbers
    import networks.helpers as RH
    it = iter(us)
    real_rand = torch.rand
    with torch.no_grad():
        torch.rand = lambda *a, **k: next(it)
        try:
            ref = RM.manipulator(pe, ve, nc, nf, ori, f_tar, args)
        finally:
            torch.rand = real_rand
        mine = O.manipulator(pc, pf, ori, list(f_tar), S, NI, args.near, args.far, labels, us=us)
    for a, b, what in zip(ref, mine, ("final_rgb", "final_ins", "tar_rgb", "tar_ins_accum")):
        assert torch.equal(a, b), "oracle != reference for manipulator %s (max diff %g)" % (what, (a - b).abs().max())

    # exchanger alone (teacher forcing): inputs as the pipeline produces them at step 1
    with torch.no_grad():
        o_raw, o_z = O.manipulator_nerf(pc, ori, None, S, args.near, args.far)
        t_raws = [O.manipulator_nerf(pc, t, None, S, args.near, args.far)[0] for t in f_tar]
        acc_o = torch.sigmoid(torch.randn(n, ins_num + 1, generator=gen) * 3)
        acc_t = [torch.sigmoid(torch.randn(n, ins_num + 1, generator=gen) * 3) for _ in tars]
        r_out = RM.exchanger(o_raw.clone(), [t.clone() for t in t_raws], acc_o.clone(), [a.clone() for a in acc_t], labels)
        m_out = O.exchanger(o_raw, t_raws, acc_o, acc_t, labels)
    assert torch.equal(r_out[0], m_out[0]) and torch.equal(r_out[2], m_out[2]) and torch.equal(r_out[3], m_out[3]), "exchanger"
    lab = m_out[2]
    print("exchanger: changed samples %d / %d; label histogram of the original: %s" %
          (int((m_out[0] != o_raw).any(-1).sum()), lab.numel(), torch.bincount(lab.reshape(-1), minlength=ins_num + 1).tolist()))
    np.savez_compressed(os.path.join(OUT, "manipulator.npz"), seed_c=31, seed_f=32, ins_num=ins_num, n_samples=S, n_importance=NI,
                        near=args.near, far=args.far, labels=np.array(labels), ins_w_c=wc["ins_linear.weight"], ins_b_c=wc["ins_linear.bias"],
                        ins_w_f=wf["ins_linear.weight"], ins_b_f=wf["ins_linear.bias"], ori=ori.numpy(), f_tar=f_tar.numpy(),
                        us=torch.stack(us, 0).numpy(), final_rgb=ref[0].numpy(), final_ins=ref[1].numpy(), tar_rgb=ref[2].numpy(),
                        tar_ins_accum=ref[3].numpy(), ex_ori_raw=o_raw.numpy(), ex_tar_raws=torch.stack(t_raws, 0).numpy(),
                        ex_acc_o=acc_o.numpy(), ex_acc_t=torch.stack(acc_t, 0).numpy(), ex_out_raw=r_out[0].numpy(),
                        ex_out_label=r_out[2].numpy(), ex_out_tar_label=r_out[3].numpy())
    print("written tests/golden/manipulator.npz", os.path.getsize(os.path.join(OUT, "manipulator.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
