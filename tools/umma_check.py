"""Quick GPU check of the tcgen05 MLP kernel against the golden fixture and the SIMT kernel, with timing.
Run on the GPU box:  python tools/umma_check.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmnerf_b200 import synth, _lib                      # noqa: E402
from dmnerf_b200.engine import get_context               # noqa: E402
from dmnerf_b200.testing import model_from_weights       # noqa: E402
from dmnerf_b200.autograd import mlp_forward_rays        # noqa: E402


def main():
    dev = "cuda"
    for ins_num in (13, 59):
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", "mlp_ins%d.npz" % ins_num)))
        net = model_from_weights(synth.make_weights(int(g["seed"]), ins_num), dev).eval()
        x = torch.from_numpy(g["x"]).to(dev)
        with torch.no_grad():
            ys = net(x, impl=_lib.IMPL_SIMT)
            try:
                yu = net(x, impl=_lib.IMPL_UMMA)
                get_context(dev).sync_check()
            except RuntimeError as e:
                print("ins_num=%d UMMA FAILED: %s" % (ins_num, e))
                continue
        ref = g["y"]
        sc = np.abs(ref).max()
        eu = np.abs(yu.cpu().numpy() - ref)
        es = np.abs(ys.cpu().numpy() - ref)
        print("ins_num=%d  scale %.3f | SIMT max err %.3e | UMMA max err %.3e (rel to scale %.3e), per-channel max: %s"
              % (ins_num, sc, es.max(), eu.max(), eu.max() / sc, np.array2string(eu.max(0)[:8], precision=2)))
        bad = np.argwhere(eu > 1e-3 * sc)
        if len(bad):
            print("   first mismatches (row, col):", bad[:10].tolist(), " rows with errors:", len(set(bad[:, 0].tolist())))
            print("   umma row0:", yu[0, :6].cpu().numpy(), "\n   ref  row0:", ref[0, :6])
    # timing on a frame-sized batch (rays mode)
    wl = synth.workload("dmsr_study")
    net = model_from_weights(synth.make_weights(202, 13), dev).eval()
    n = 307200 // 4
    ro, rd = torch.from_numpy(wl["rays_o"][:n]).to(dev), torch.from_numpy(wl["rays_d"][:n]).to(dev)
    z = (torch.rand(n, 192, device=dev).sort(-1).values * 11 + 4).contiguous()
    for impl, name in ((_lib.IMPL_UMMA, "UMMA"), (_lib.IMPL_SIMT, "SIMT")):
        try:
            with torch.no_grad():
                o = mlp_forward_rays(net, ro, rd, z, impl)
                get_context(dev).sync_check()
                t0 = time.perf_counter()
                reps = 3 if impl == _lib.IMPL_UMMA else 1
                for _ in range(reps):
                    o = mlp_forward_rays(net, ro, rd, z, impl)
                get_context(dev).sync_check()
                dt = (time.perf_counter() - t0) / reps
            fl = 2.0 * synth.macs_per_sample(13) * n * 192
            print("%s fine MLP on %d rays x 192: %.2f ms -> %.1f algorithmic TFLOP/s, %.0f rays/s-equivalent (fine only)"
                  % (name, n, dt * 1e3, fl / dt / 1e12, n / dt))
            if impl == _lib.IMPL_UMMA:
                ou = o
            else:
                d = (ou - o).abs().max().item()
                print("UMMA vs SIMT on the big batch: max abs diff %.3e (scale %.2f)" % (d, o.abs().max().item()))
        except RuntimeError as e:
            print(name, "FAILED:", e)


if __name__ == "__main__":
    main()
