"""Seeded synthetic workloads for the DM-NeRF render path (no dataset / checkpoint is shipped).

Everything here is numpy-only and deterministic across machines (PCG64 streams), so that the
golden fixtures under tests/golden/, the GPU parity tests, bench.py and smoke() all see the same
weights and rays without storing multi-MB tensors in git.

Workload definitions follow SURVEY.md section 8(d):
  * weights: nn.Linear default init U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for the layer list of
    reference networks/dm_nerf.py:59-78, optionally "trained-like" (density_linear.weight x30) so
    rays actually terminate inside [near, far] instead of on the 1e10 tail sample.
  * rays: pin-hole cameras as built by reference networks/helpers.py:50-61 (get_rays_k) with the
    DM-SR intrinsics of datasets/loader_dmsr.py:136-137 or the Replica intrinsics of
    datasets/loader_replica.py:93-94.
"""
import math

import numpy as np

# (name, out_features, in_features) in reference state_dict order (networks/dm_nerf.py:65-78)
def layer_table(ins_num, W=256, D=8, ch_pts=63, ch_views=27, skips=(4,)):
    rows = [("mlps.0", W, ch_pts)]
    for i in range(D - 1):
        rows.append(("mlps.%d" % (i + 1), W, W + ch_pts if i in skips else W))
    rows += [
        ("rgb_feature_linear", W, W),
        ("ins_feature_linear", W, W),
        ("rgb_feature_linears.0", W // 2, W + ch_views),
        ("ins_feature_linears.0", W // 2, W),
        ("density_linear", 1, W),
        ("ins_linear", ins_num + 1, W // 2),
        ("rgb_linear", 3, W // 2),
    ]
    return rows


def param_names(ins_num=13):
    names = []
    for n, _, _ in layer_table(ins_num):
        names += [n + ".weight", n + ".bias"]
    return names


def make_weights(seed, ins_num=13, trained_like=True, density_gain=30.0):
    """Return {state_dict key: float32 ndarray} for one DM_NeRF(8,256,63,27,[4],ins_num)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for name, fo, fi in layer_table(ins_num):
        bound = 1.0 / math.sqrt(fi)
        out[name + ".weight"] = rng.uniform(-bound, bound, size=(fo, fi)).astype(np.float32)
        out[name + ".bias"] = rng.uniform(-bound, bound, size=(fo,)).astype(np.float32)
    if trained_like:
        out["density_linear.weight"] = (out["density_linear.weight"] * np.float32(density_gain)).astype(np.float32)
    return out


def macs_per_sample(ins_num=13):
    return sum(fo * fi for _, fo, fi in layer_table(ins_num))


def flops_per_ray(ins_num=13, n_coarse=64, n_importance=128):
    """Algorithmic forward FLOPs per ray: 2 * (S + S+I) * MACs/sample."""
    return 2.0 * (n_coarse + n_coarse + n_importance) * macs_per_sample(ins_num)


def algorithmic_bytes_per_ray(ins_num=13):
    """24 B in (o,d) + rgb 12 + depth 4 + acc 4 + 4*ins_num out."""
    return 24 + 12 + 4 + 4 + 4 * ins_num


def rot_phi(phi):
    c, s = math.cos(phi), math.sin(phi)
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], dtype=np.float64)


def rot_theta(th):
    c, s = math.cos(th), math.sin(th)
    return np.array([[c, 0, -s, 0], [0, 1, 0, 0], [s, 0, c, 0], [0, 0, 0, 1]], dtype=np.float64)


def pose_spherical(theta_deg, phi_deg, radius):
    """Spherical camera-to-world pose (same convention as reference tools/pose_generator.py:29-34)."""
    t = np.eye(4)
    t[2, 3] = radius
    c2w = rot_theta(theta_deg / 180.0 * math.pi) @ rot_phi(phi_deg / 180.0 * math.pi) @ t
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float64)
    return (flip @ c2w).astype(np.float32)


def dmsr_intrinsics(H=480, W=640, fovx=0.6911):
    f = 0.5 * W / math.tan(0.5 * fovx)
    return np.array([[f, 0, W / 2.0], [0, -f, H / 2.0], [0, 0, -1]], dtype=np.float32)


def replica_intrinsics(H=480, W=640):
    return np.array([[W / 2.0, 0, (W - 1) / 2.0], [0, W / 2.0, (H - 1) / 2.0], [0, 0, 1]], dtype=np.float32)


def rays_from_camera(H, W, K, c2w):
    """numpy twin of reference networks/helpers.py:50-61 (get_rays_k); returns [H*W,3] o and d."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    dirs = np.stack([(i - K[0, 2]) / K[0, 0], (j - K[1, 2]) / K[1, 1], K[2, 2] * np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., None, :] * c2w[:3, :3], -1).astype(np.float32)
    rays_o = np.broadcast_to(c2w[:3, -1], rays_d.shape).astype(np.float32)
    return rays_o.reshape(-1, 3).copy(), rays_d.reshape(-1, 3).copy()


WORKLOADS = {
    # name: (H, W, intrinsics fn, near, far, ins_num, pose)
    "dmsr_study": dict(H=480, W=640, K="dmsr", near=4.0, far=15.0, ins_num=13, pose=(30.0, -65.0, 7.0)),
    "replica_room0": dict(H=480, W=640, K="replica", near=0.0, far=6.5, ins_num=59, pose=(10.0, -20.0, 1.5)),
    "replica_room0_93": dict(H=480, W=640, K="replica", near=0.0, far=6.5, ins_num=93, pose=(10.0, -20.0, 1.5)),
    "replica_office2": dict(H=480, W=640, K="replica", near=0.0, far=5.7, ins_num=69, pose=(0.0, -15.0, 1.2)),
}


def workload(name, frame=0):
    """Return dict(rays_o, rays_d, near, far, ins_num, H, W) for one synthetic frame of a workload."""
    w = WORKLOADS[name]
    K = dmsr_intrinsics(w["H"], w["W"]) if w["K"] == "dmsr" else replica_intrinsics(w["H"], w["W"])
    th, ph, r = w["pose"]
    c2w = pose_spherical(th + 0.4 * frame, ph, r)   # smooth synthetic trajectory
    o, d = rays_from_camera(w["H"], w["W"], K, c2w)
    return dict(rays_o=o, rays_d=d, near=w["near"], far=w["far"], ins_num=w["ins_num"],
                H=w["H"], W=w["W"], K=K, c2w=c2w)
