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
