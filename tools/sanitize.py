"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmnerf_b200 import synth, _lib                      # noqa: E402
from dmnerf_b200.engine import get_context               # noqa: E402
from dmnerf_b200.testing import make_models              # noqa: E402
from dmnerf_b200.render import render_rays, dm_nerf      # noqa: E402
from dmnerf_b200.backward import render_rays_grad        # noqa: E402
from dmnerf_b200.helpers import get_rays_k, sample_pdf   # noqa: E402

dev = "cuda"
wl = synth.workload("dmsr_study")
nc, nf, _, _ = make_models(1, 2, 13, dev)
n = 37
sel = np.linspace(0, 307199, n).astype(np.int64)
ro, rd = torch.from_numpy(wl["rays_o"][sel]).to(dev), torch.from_numpy(wl["rays_d"][sel]).to(dev)
z = (torch.linspace(0, 1, 64) * 11 + 4).to(dev)
with torch.no_grad():
    a = render_rays(ro, rd, nc, nf, z, want_raw=False, want_samples=True)          # fused kernel, odd ray count
    b = render_rays(ro, rd, nc, nf, z, want_raw=True, impl=_lib.IMPL_UMMA)         # unfused tcgen05 + stage kernels
    c = render_rays(ro, rd, nc, nf, z, want_raw=True, impl=_lib.IMPL_SIMT)         # fp32 kernels
    o, d = get_rays_k(48, 64, wl["K"], torch.from_numpy(wl["c2w"]).to(dev))
    s = sample_pdf(torch.sort(torch.rand(5, 63, device=dev)).values, torch.rand(5, 62, device=dev), 128, det=False)
nc.train(); nf.train()
out = render_rays_grad(ro, rd, nc, nf, z, perturb=1.0)
(out["rgb_fine"].sum() + out["ins_fine"].sum() + out["raw_coarse"].sum() * 1e-3).backward()
# round-2 kernels: ray selection, Hungarian-matched loss, penalizer (both tile sizes), exchanger-free edit primitives
import types                                             # noqa: E402
from dmnerf_b200.helpers import get_select_full          # noqa: E402
from dmnerf_b200.evaluator import ins_criterion          # noqa: E402
from dmnerf_b200.penalizer import ins_penalizer          # noqa: E402
img = torch.rand(48, 64, 3, device=dev)
lab = (torch.arange(48 * 64, device=dev).reshape(48, 64) % 5).to(torch.int16)
np.random.seed(0)
tc, ti, rays = get_select_full(img, torch.from_numpy(wl["c2w"]).to(dev), synth.dmsr_intrinsics(48, 64), lab, 64)
pred = torch.sigmoid(torch.randn(64, 13, device=dev)).requires_grad_(True)
ins_criterion(pred, ti, 13)[0].sum().backward()
pargs = types.SimpleNamespace(tolerance=0.05, deta_w=0.05)
for cc in (18, 64, 132):
    praw = torch.randn(9, 50, cc, device=dev, requires_grad=True)
    pz = torch.rand(9, 50, device=dev).sort(-1).values * 11 + 4
    ins_penalizer(praw, pz, pz[:, 20].clone(), torch.randn(9, 3, device=dev), pargs).sum().backward()
# late round 2: device-side pixel selection, host-side assignment path, host entry point rendered in parts (second stream)
from dmnerf_b200.helpers import select_pixels            # noqa: E402
pix = select_pixels(48, 64, 48 * 64, dev, seed=3)
assert sorted(pix.tolist()) == list(range(48 * 64))
os.environ["DMNERF_INS_ASSIGN"] = "host"
ins_criterion(pred, ti, 13)[0].sum().backward()
del os.environ["DMNERF_INS_ASSIGN"]
if os.environ.get("SANITIZE_PARTS", "1") == "1":
    nn = 131074
    hro = torch.from_numpy(wl["rays_o"][:nn]).contiguous().pin_memory()
    hrd = torch.from_numpy(wl["rays_d"][:nn]).contiguous().pin_memory()
    hz = z.cpu().contiguous().pin_memory()
    hout = torch.empty(nn, 3).pin_memory()
    io = _lib.RenderIO()
    io.rays_o, io.rays_d, io.z_coarse, io.z_row_stride, io.rgb_fine = _lib.ptr(hro), _lib.ptr(hrd), _lib.ptr(hz), 0, _lib.ptr(hout)
    nc.eval(); nf.eval()
    cx = get_context(torch.device(dev))
    cx.bind(0, nc); cx.bind(1, nf)
    _lib.check(cx.lib.dmnerf_render_forward_host(cx.handle, io, nn, 64, 128, 0, 0, cx.stream()), "dmnerf_render_forward_host")
    assert torch.isfinite(hout).all()
get_context(dev).sync_check()
print("sanitize run ok", float(a["rgb_fine"].sum()), float(b["rgb_fine"].sum()), float(c["rgb_fine"].sum()))
