"""Generate tests/golden/penalizer.npz from the UNMODIFIED reference (networks/penalizer.py) and pin the oracle to it
(value and gradient, bit for bit on the CPU).      python oracle/make_golden_penalizer.py"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from networks.penalizer import ins_penalizer as ref_ins_penalizer      # noqa: E402
from oracle import dmnerf_oracle as O                                    # noqa: E402

torch.autograd.set_detect_anomaly(False)
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    gen = torch.Generator().manual_seed(77)
    save = {}
    for tag, n, s, k in (("a", 24, 64, 14), ("b", 12, 192, 60)):
        raw = (torch.randn(n, s, 4 + k, generator=gen) * 2.0)
        z = (torch.rand(n, s, generator=gen).sort(-1).values * 11 + 4)
        rd = torch.randn(n, 3, generator=gen) * 1.3
        # rendered depths: inside the sampled range for most rays, outside for a few (all-before / all-after masks)
        depth = z[torch.arange(n), torch.randint(0, s, (n,), generator=gen)] + 0.01
        depth[0] = 0.5
        depth[1] = 100.0
        args = types.SimpleNamespace(tolerance=0.05, deta_w=0.05)
        r1 = raw.clone().requires_grad_(True)
        ref = ref_ins_penalizer(r1, z, depth, rd, args)
        ref.sum().backward()
        r2 = raw.clone().requires_grad_(True)
        mine = O.ins_penalizer(r2, z, depth, rd, args.tolerance, args.deta_w)
        mine.sum().backward()
        assert torch.equal(ref, mine) and torch.equal(r1.grad, r2.grad), "oracle != reference (penalizer %s)" % tag
        save.update({"raw_" + tag: raw.numpy(), "z_" + tag: z.numpy(), "rays_d_" + tag: rd.numpy(), "depth_" + tag: depth.numpy(),
                     "loss_" + tag: ref.detach().numpy(), "grad_" + tag: r1.grad.numpy()})
        print("penalizer case %s: loss %.7g, |grad| max %.3g" % (tag, float(ref), float(r1.grad.abs().max())))
    np.savez_compressed(os.path.join(OUT, "penalizer.npz"), tolerance=0.05, deta_w=0.05, **save)
    print("written", os.path.join(OUT, "penalizer.npz"), os.path.getsize(os.path.join(OUT, "penalizer.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
