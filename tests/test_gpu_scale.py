"""GPU: the TIMED kernel (the fused render kernel, render_rays(want_raw=False) = mlp_umma_kernel<true>) against the LIVE CPU
oracle on 2048 strided rays of every BASELINE config (ins_num 13 / 59 / 93 / 69), with the error distribution of every map
and the oracle's own fp64 twin as the yard-stick for the ill-conditioned fine pass (sample_pdf amplifies last-bit differences
of the coarse weights: the reference moves as much when its own arithmetic is carried out in fp64).  No ray is excluded."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from dmnerf_b200 import _lib
from dmnerf_b200.testing import format_parity_table

N_RAYS = 2048


@pytest.mark.parametrize("name", ["dmsr_study", "replica_room0", "replica_room0_93", "replica_office2"])
def test_fused_kernel_vs_live_oracle_error_distribution(name):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from parity_at_scale import run_config
    table, info = run_config(name, N_RAYS)
    assert info["render_launches"] == 1, "the fused path must be ONE kernel launch"
    print("\n" + format_parity_table("%s (ins_num %d, %d rays)" % (name, info["ins_num"], N_RAYS), table))
    for k, row in table.items():
        o, t = row["ours"], row["twin"]
        assert o["n"] == N_RAYS
        if k.endswith("_coarse"):
            # no sampling in front of the coarse maps: every ray inside the north_star tolerance, tails at the 1e-5 level
            assert o["frac_within"] >= 0.999 and o["max"] <= 5e-5, (k, o)
        else:
            # fine maps: as close to the fp32 reference as the reference's own fp64 twin is
            # (the tails -- max, and through it the PSNR -- hang on a handful of rays whose importance samples hop a bin, in
            #  either arithmetic: they get the wider factors; measured: max <= 2.6 x, PSNR >= twin - 5.0 dB)
            assert o["frac_within"] >= t["frac_within"] - 0.03, (k, o, t)
            assert o["median"] <= 2.0 * t["median"] + 1e-6, (k, o, t)
            assert o["p99"] <= 2.0 * t["p99"] + 1e-5, (k, o, t)
            assert o["max"] <= 5.0 * t["max"] + 1e-4, (k, o, t)
            assert o["psnr"] >= t["psnr"] - 8.0, (k, o, t)
            assert o["psnr"] >= 55.0, (k, o)
