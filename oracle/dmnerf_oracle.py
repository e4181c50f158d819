"""CPU oracle for the DM-NeRF per-ray render path -- TEST INFRASTRUCTURE ONLY.

This file is a functional, dtype-generic restatement (torch CPU ops, fp32 by default, fp64 for the
noise-floor twin) of the reference algorithm in /root/reference/networks/{render,dm_nerf,helpers}.py.
It exists only to check the CUDA path: only tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py may import it.  The product package (dm-nerf_b200/) never does.

Parity pinning: the reference ships NO tests, golden vectors or checkpoints, so
the pin is the reference code itself: oracle/make_golden.py imports the unmodified reference from
/root/reference in the build container, checks every function below against it on seeded inputs
and writes the reference's outputs to tests/golden/*.npz.  tests/test_oracle.py re-checks this oracle against
those committed fixtures wherever the tests run.

Every function cites the reference lines it restates.  Weights are passed as a plain dict keyed by
the reference state_dict names (networks/dm_nerf.py:65-78).
"""
import numpy as np
import torch
import torch.nn.functional as F

SKIPS = (4,)
N_TRUNK = 8


def embed(x, n_freqs):
    """Positional encoding, reference networks/dm_nerf.py:13-38 with get_embedder's kwargs (:45-52):
    [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]; frequencies are exact powers of two."""
    parts = [x]
    for k in range(n_freqs):
        xf = x * float(2 ** k)
        parts.append(torch.sin(xf))
        parts.append(torch.cos(xf))
    return torch.cat(parts, -1)


def _lin(p, name, x):
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def mlp_forward(p, x, ch_pts=63, ch_views=27):
    """DM_NeRF.forward, reference networks/dm_nerf.py:80-106. x [M, 63+27] -> [M, 3+1+ins_num+1]."""
    pts, dirs = x[..., :ch_pts], x[..., ch_pts:ch_pts + ch_views]
    h = pts
    for i in range(N_TRUNK):
        h = torch.relu(_lin(p, "mlps.%d" % i, h))                                  # :84-85
        if i in SKIPS:
            h = torch.cat([h, pts], -1)                                            # :86-87  order [h, pts]
    rgb_f = _lin(p, "rgb_feature_linear", h)                                        # :89 (no activation)
    rgb_f = torch.relu(_lin(p, "rgb_feature_linears.0", torch.cat([rgb_f, dirs], -1)))   # :90-93
    ins_f = _lin(p, "ins_feature_linear", h.detach())                               # :95-96 (h is detached)
    ins_f = torch.relu(_lin(p, "ins_feature_linears.0", ins_f))                     # :97-99
    density = _lin(p, "density_linear", h)                                          # :101
    rgb = _lin(p, "rgb_linear", rgb_f)                                              # :102
    ins = _lin(p, "ins_linear", ins_f)                                              # :103
    return torch.cat([rgb, density, ins], -1)                                       # :105


def composite(raw, z_vals, rays_d, keep_all_ins=False):
    """render_train, reference networks/render.py:6-28 (keep_all_ins=True gives manipulator_render,
    networks/manipulator.py:86-105).  Returns rgb_map, weights, depth_map, ins_map, acc_map."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    tail = torch.full_like(dists[..., :1], 1e10)                                    # :10
    dists = torch.cat([dists, tail], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)                        # :12
    rgb = torch.sigmoid(raw[..., :3])                                               # :14
    alpha = 1.0 - torch.exp(-torch.relu(raw[..., 3]) * dists)                       # :7,16
    ones = torch.ones_like(alpha[..., :1])
    trans = torch.cumprod(torch.cat([ones, 1.0 - alpha + 1e-10], -1), -1)[..., :-1]  # :18
    weights = alpha * trans
    rgb_map = torch.sum(weights[..., None] * rgb, -2)                               # :19
    depth_map = torch.sum(weights * z_vals, -1)                                     # :20
    w_ins = weights if keep_all_ins else weights.detach()                           # :22-23 (manipulator.py:100 keeps grad)
    ins_map = torch.sigmoid(torch.sum(w_ins[..., None] * raw[..., 4:], -2))         # :24-25
    if not keep_all_ins:
        ins_map = ins_map[..., :-1]                                                 # :26
    acc_map = torch.sum(weights, -1)          # not in the reference; SURVEY.md name-mapping table
    return rgb_map, weights, depth_map, ins_map, acc_map


def sample_pdf(bins, weights, n_samples, det=False, u=None):
    """Inverse-CDF sampling, reference networks/helpers.py:123-155.  `u` replaces the reference's
    internal torch.rand draw (:135) so the caller controls the random stream."""
    weights = weights + 1e-5                                                        # :125
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)                      # :128
    if det:
        u = torch.linspace(0.0, 1.0, steps=n_samples, dtype=bins.dtype)             # :132
        u = u.expand(list(cdf.shape[:-1]) + [n_samples])
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)                                   # :139
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bin_b, bin_a = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)                # :151
    t = (u - cdf_b) / denom
    return bin_b + t * (bin_a - bin_b)                                              # :153


def z_val_sample(n_rays, near, far, n_samples, dtype=torch.float32):
    """reference networks/helpers.py:114-119: near + linspace(0,1,S) * (far - near), expanded to N rows."""
    t = torch.linspace(0.0, 1.0, steps=n_samples, dtype=dtype)
    near_t = near * torch.ones((n_rays, 1), dtype=dtype)
    far_t = far * torch.ones((n_rays, 1), dtype=dtype)
    return (near_t + t * (far_t - near_t)).expand(n_rays, n_samples)


def get_rays_k(H, W, K, c2w):
    """reference networks/helpers.py:50-61."""
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - K[0, 2]) / K[0, 0], (j - K[1, 2]) / K[1, 1], K[2, 2] * torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def stratify(z, t_rand):
    """Stratified jitter, reference networks/render.py:40-47."""
    mids = 0.5 * (z[..., 1:] + z[..., :-1])
    upper = torch.cat([mids, z[..., -1:]], -1)
    lower = torch.cat([z[..., :1], mids], -1)
    return lower + (upper - lower) * t_rand


def _net_inputs(rays_o, rays_d, viewdirs, z):
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]             # render.py:49,71
    flat = pts.reshape(-1, 3)
    dirs = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)                       # render.py:55-56
    return torch.cat([embed(flat, 10), embed(dirs, 4)], -1), pts.shape[:-1]          # render.py:54-58


def render(rays_o, rays_d, p_coarse, p_fine, z_coarse, perturb=0.0, n_importance=128,
           t_rand=None, u=None, is_train=False, n_ins=None):
    """dm_nerf(), reference networks/render.py:31-96.  t_rand [N,S] and u [N,I] are the two uniform
    draws the reference takes from torch.rand (render.py:46, helpers.py:135), in that order."""
    viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)                    # :37
    if perturb > 0.0:
        z_coarse = stratify(z_coarse, t_rand)                                       # :40-47
    x, shp = _net_inputs(rays_o, rays_d, viewdirs, z_coarse)
    raw_c = mlp_forward(p_coarse, x).reshape(*shp, -1)                              # :60-61
    rgb_c, w_c, depth_c, ins_c, acc_c = composite(raw_c, z_coarse, rays_d)           # :63
    z_mid = 0.5 * (z_coarse[..., 1:] + z_coarse[..., :-1])                          # :66
    z_samples = sample_pdf(z_mid, w_c[..., 1:-1], n_importance, det=(perturb == 0.0), u=u).detach()  # :67-68
    z_fine, _ = torch.sort(torch.cat([z_coarse, z_samples], -1), -1)                # :70
    x, shp = _net_inputs(rays_o, rays_d, viewdirs, z_fine)
    raw_f = mlp_forward(p_fine, x).reshape(*shp, -1)                                # :82-83
    rgb_f, w_f, depth_f, ins_f, acc_f = composite(raw_f, z_fine, rays_d)             # :86
    if is_train and n_ins is not None:
        ins_f, ins_c = ins_f[-n_ins:], ins_c[-n_ins:]                               # :88-90
    return {"rgb_fine": rgb_f, "ins_fine": ins_f, "z_vals_fine": z_fine, "raw_fine": raw_f,
            "raw_coarse": raw_c, "rgb_coarse": rgb_c, "ins_coarse": ins_c, "z_vals_coarse": z_coarse,
            "depth_fine": depth_f, "depth_coarse": depth_c,
            # extras (not in the reference dict): weights and acc maps
            "weights_coarse": w_c, "weights_fine": w_f, "acc_coarse": acc_c, "acc_fine": acc_f}


def to_torch(weights_np, dtype=torch.float32):
    return {k: torch.from_numpy(v).to(dtype) for k, v in weights_np.items()}


def train_loss(out, target_rgb):
    """Scalar loss used for the training-step parity case C4: the rgb MSE terms of
    train_dmsr.py:35-37 plus mean(ins) stand-ins for the instance terms (the Hungarian loss itself is
    out of scope)."""
    return (((out["rgb_coarse"] - target_rgb) ** 2).mean() + ((out["rgb_fine"] - target_rgb) ** 2).mean()
            + out["ins_coarse"].mean() + out["ins_fine"].mean())


def emptiness_penalizer(raw, z_vals, depths, rays_d, tolerance, deta_w):
    """"Emptiness" regulariser on the per-sample object logits, reference networks/penalizer.py:5-55.
    raw [N,S,4+K], z_vals [N,S], depths [N,1] (already detached, penalizer.py:59), rays_d [N,3] -> loss tensor of shape [1].
    In front of the surface (more than `tolerance` before the rendered depth) every sample should be "no object" (last class),
    weighted by 1 - gaussian(distance to the surface); inside the +-tolerance shell the last class is pushed to 0, weighted by
    the gaussian.  Both terms are masked means (penalizer.py:41-42, 52)."""
    sigma_h = torch.Tensor([0.4])                                                   # :10
    sigma_w = torch.Tensor([deta_w])
    two_pi_root = torch.sqrt(torch.Tensor([2 * np.pi]))

    def gaussian(delta):                                                            # :7-8
        return torch.exp(-(delta ** 2) / (2 * (sigma_w ** 2))) / (sigma_h * two_pi_root) + 1e-8

    norm = torch.norm(rays_d[..., None, :], dim=-1)                                 # :13   [N,1]
    front = (depths - tolerance) * norm                                             # :14,16
    back = (depths + tolerance) * norm                                              # :15,17
    surface = depths * norm                                                         # :18
    pos = z_vals * norm                                                             # :19
    g = gaussian(surface - pos)                                                     # :22-23
    air = 1 - g                                                                     # :24
    m_before = (pos < front).type(torch.float32)                                    # :27
    m_after = (pos > back).type(torch.float32)                                      # :28
    m_middle = 1 - (m_after + m_before)                                             # :29
    pred = torch.sigmoid(raw[..., 4:])                                              # :32-33
    gt = torch.zeros_like(pred)
    gt[..., -1] = 1                                                                 # :37-38
    l_before = -gt * torch.log(pred + 1e-8) - (1 - gt) * torch.log(1 - pred + 1e-8)  # :39
    l_before = l_before * (air * m_before)[..., None]                               # :40-41
    l_before = torch.sum(l_before) / (pred.shape[-1] * torch.maximum(torch.sum(m_before), torch.tensor([1e-8])))   # :42-43
    last = pred[..., -1]                                                            # :46
    gt_mid = torch.zeros_like(last)
    l_mid = -gt_mid * torch.log(last + 1e-8) - (1 - gt_mid) * torch.log(1 - last + 1e-8)   # :48-49
    l_mid = l_mid * (g * m_middle)                                                  # :50-51
    l_mid = torch.sum(l_mid) / torch.maximum(torch.sum(m_middle), torch.tensor([1e-8]))    # :52
    return l_before + l_mid                                                         # :53


def ins_penalizer(raw, z_vals, depth, rays_d, tolerance, deta_w):
    """networks/penalizer.py:58-62 (args.tolerance / args.deta_w passed explicitly)."""
    return emptiness_penalizer(raw, z_vals, depth[..., None].detach(), rays_d, tolerance, deta_w)
