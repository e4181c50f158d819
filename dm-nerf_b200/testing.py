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
