"""Helpers shared by tests, smoke() and bench.py (product side; does not import oracle/)."""
import numpy as np
import torch

from . import synth
from .model import DM_NeRF


def model_from_weights(weights_np, device="cuda"):
    ins_num = weights_np["ins_linear.weight"].shape[0] - 1
    net = DM_NeRF(8, 256, 63, 27, [4], ins_num)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in weights_np.items()})
    return net.to(device)


def make_models(seed_coarse, seed_fine, ins_num, device="cuda", trained_like=True):
    wc = synth.make_weights(seed_coarse, ins_num, trained_like)
    wf = synth.make_weights(seed_fine, ins_num, trained_like)
    return model_from_weights(wc, device), model_from_weights(wf, device), wc, wf


def max_rel_err(got, ref, abs_floor):
    """max |got-ref| / max(|ref|, abs_floor) -- the stage-wise parity metric."""
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), abs_floor)))


def frac_bad(got, ref, rtol, atol):
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    bad = np.abs(got - ref) > (atol + rtol * np.abs(ref))
    return float(bad.mean())


def scale_err(got, ref):
    """max |got-ref| / max |ref|: error relative to the tensor's own scale (outputs of a GEMM chain are sums with
    cancellation, so an element-wise relative error is undefined near zero)."""
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-30))


def rel_l2(got, ref):
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    return float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30))


def raw_errs(got, ref):
    """Parity metrics of a network output [..., C] per channel group (rgb logits | density | instance logits): the groups
    have different scales (trained-like density weights are x30)."""
    got = np.asarray(got); ref = np.asarray(ref)
    groups = (slice(0, 3), slice(3, 4), slice(4, None))
    return max(scale_err(got[..., g], ref[..., g]) for g in groups), max(rel_l2(got[..., g], ref[..., g]) for g in groups)


# ---- error-distribution statistics of a rendered map against the reference (parity at scale: tests/test_gpu_scale.py,
#      tools/parity_at_scale.py, smoke()) ----------------------------------------------------------------------------------
MAP_KEYS = ("rgb_coarse", "depth_coarse", "acc_coarse", "ins_coarse", "rgb_fine", "depth_fine", "acc_fine", "ins_fine")


def error_stats(got, ref, rtol=1e-4):
    """Per-RAY error distribution of a map [N] or [N,C] against the reference.  A value is "within rtol" when
    |got - ref| <= rtol * max(|ref|, 0.1 * scale), scale = max |ref| over the map (values far below the map's scale are
    judged against a tenth of it, not against themselves); a ray is within when all its channels are.  Errors are reported
    relative to the scale: median / p99 / max over rays of the per-ray max error; psnr = 10 log10(scale^2 / mse)."""
    got = np.asarray(got, dtype=np.float64).reshape(len(got), -1)
    ref = np.asarray(ref, dtype=np.float64).reshape(len(ref), -1)
    scale = max(float(np.abs(ref).max()), 1e-30)
    err = np.abs(got - ref)
    ok = (err <= rtol * np.maximum(np.abs(ref), 0.1 * scale)).all(axis=1)
    per_ray = err.max(axis=1) / scale
    mse = float(np.mean((got - ref) ** 2))
    return {"frac_within": float(ok.mean()), "median": float(np.median(per_ray)), "p99": float(np.quantile(per_ray, 0.99)),
            "max": float(per_ray.max()), "psnr": float(10.0 * np.log10(scale * scale / max(mse, 1e-300))), "n": int(len(got))}


def parity_table(ours, twin, ref, keys=MAP_KEYS):
    """{key: {"ours": stats, "twin": stats}}: our maps and the reference's own fp64 twin, both against the fp32 reference."""
    to_np = lambda t: t.detach().cpu().double().numpy() if hasattr(t, "detach") else np.asarray(t)
    return {k: {"ours": error_stats(to_np(ours[k]), to_np(ref[k])), "twin": error_stats(to_np(twin[k]), to_np(ref[k]))}
            for k in keys if k in ours and k in twin and k in ref}


def format_parity_table(name, table):
    lines = ["| %s | within 1e-4 (ours / fp64 twin) | median | p99 | max | PSNR dB |" % name, "|---|---|---|---|---|---|"]
    for k, row in table.items():
        o, t = row["ours"], row["twin"]
        lines.append("| %s | %.4f / %.4f | %.1e / %.1e | %.1e / %.1e | %.1e / %.1e | %.1f / %.1f |" %
                     (k, o["frac_within"], t["frac_within"], o["median"], t["median"], o["p99"], t["p99"], o["max"], t["max"],
                      o["psnr"], t["psnr"]))
    return "\n".join(lines)
