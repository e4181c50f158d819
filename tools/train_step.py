"""Minimal driver for ncu: a few training steps (BASELINE configs[3]: 1024 rays, forward + backward + Adam).
    python tools/train_step.py [steps]"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmnerf_b200 import synth                              # noqa: E402
from dmnerf_b200.testing import make_models                # noqa: E402
from dmnerf_b200.render import dm_nerf                     # noqa: E402
from dmnerf_b200.embedder import get_embedder              # noqa: E402
from dmnerf_b200.engine import get_context                 # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = "cuda"
wl = synth.workload("dmsr_study")
nc, nf, _, _ = make_models(201, 202, 13, dev)
sel = np.random.Generator(np.random.PCG64(0)).choice(307200, 1024, replace=False)
ro, rd = torch.from_numpy(wl["rays_o"][sel]).to(dev), torch.from_numpy(wl["rays_d"][sel]).to(dev)
rays = torch.stack([ro, rd], 0)
targs = types.SimpleNamespace(perturb=1.0, N_importance=128, is_train=True, N_ins=None)
pe, ve = get_embedder(10)[0], get_embedder(4)[0]
zc = torch.linspace(float(wl["near"]), float(wl["far"]), 64, device=dev)[None].expand(1024, 64)
tgt = torch.rand(1024, 3, device=dev)
nc.train(); nf.train()
opt = torch.optim.Adam(list(nc.parameters()) + list(nf.parameters()), lr=5e-4)
for _ in range(steps):
    out = dm_nerf(rays, pe, ve, nc, nf, zc, targs)
    loss = ((out["rgb_fine"] - tgt) ** 2).mean() + ((out["rgb_coarse"] - tgt) ** 2).mean() + out["ins_fine"].mean()
    opt.zero_grad()
    loss.backward()
    opt.step()
get_context(dev).sync_check()
print("done", steps, float(loss))
