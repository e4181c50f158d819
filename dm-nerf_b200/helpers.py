"""Drop-in for the hot-path functions of reference networks/helpers.py: sample_pdf (:123-155),
z_val_sample (:114-119), get_rays_k (:50-61)."""
import torch

from . import _lib
from .engine import get_context


def sample_pdf(bins, weights, N_samples, det=False, u=None):
    """Inverse-CDF sampling on the GPU.  `u` (optional, [N, N_samples]) overrides the internal
    torch.rand draw that the reference makes when det is False (helpers.py:135)."""
    if not bins.is_cuda:
        raise RuntimeError("sample_pdf: expected CUDA tensors (no CPU fallback)")
    lead = bins.shape[:-1]
    nb = bins.shape[-1]
    b = bins.reshape(-1, nb).contiguous().float()
    w = weights.reshape(-1, nb - 1).contiguous().float()
    n = b.shape[0]
    if det:
        u_t = None
    else:
        u_t = (torch.rand(list(lead) + [N_samples], device=bins.device) if u is None else u)
        u_t = u_t.reshape(-1, N_samples).contiguous().float()
    out = torch.empty((n, N_samples), device=bins.device, dtype=torch.float32)
    ctx = get_context(bins.device)
    _lib.check(ctx.lib.dmnerf_sample_pdf(_lib.ptr(b), _lib.ptr(w), n, nb, N_samples, _lib.ptr(u_t), _lib.ptr(out),
                                         ctx.stream()), "dmnerf_sample_pdf")
    return out.reshape(*lead, N_samples)


def sort_concat(a, b):
    """torch.sort(torch.cat([a, b], -1), -1).values (render.py:70) as one kernel."""
    lead = a.shape[:-1]
    a2, b2 = a.reshape(-1, a.shape[-1]).contiguous().float(), b.reshape(-1, b.shape[-1]).contiguous().float()
    out = torch.empty((a2.shape[0], a2.shape[1] + b2.shape[1]), device=a.device, dtype=torch.float32)
    ctx = get_context(a.device)
    _lib.check(ctx.lib.dmnerf_sort_concat(_lib.ptr(a2), _lib.ptr(b2), a2.shape[0], a2.shape[1], b2.shape[1],
                                          _lib.ptr(out), ctx.stream()), "dmnerf_sort_concat")
    return out.reshape(*lead, -1)


def z_val_sample(N_rays, near, far, N_samples, device=None):
    """near + linspace(0,1,S)*(far-near), expanded (stride 0) to N_rays rows like the reference."""
    near_t = near * torch.ones(size=(N_rays, 1), device=device)
    far_t = far * torch.ones(size=(N_rays, 1), device=device)
    t_vals = torch.linspace(0., 1., steps=N_samples, device=device)
    return (near_t[:1] + t_vals * (far_t[:1] - near_t[:1])).expand([N_rays, N_samples])


def get_rays_k(H, W, K, c2w):
    """Camera rays for intrinsics K and pose c2w, same arithmetic as the reference (helpers.py:50-61)."""
    K = torch.as_tensor(K, dtype=torch.float32, device=c2w.device)
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=c2w.device),
                          torch.linspace(0, H - 1, H, device=c2w.device), indexing="ij")
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - K[0, 2]) / K[0, 0], (j - K[1, 2]) / K[1, 1], K[2, 2] * torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d
