"""Generate tests/golden/evaluator.npz from the UNMODIFIED reference (networks/evaluator.py: ins_criterion / hungarian) and pin
the oracle to it (loss parts, cost matrices, assignment and gradient, bit for bit on the CPU).
    python oracle/make_golden_evaluator.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from networks.evaluator import ins_criterion as ref_ins_criterion, hungarian as ref_hungarian      # noqa: E402
from oracle import dmnerf_oracle as O                                                                 # noqa: E402

torch.autograd.set_detect_anomaly(False)
OUT = os.path.join(ROOT, "tests", "golden")


def case(gen, n, k, labels_present):
    """Rendered instance maps look like sigmoid(logits) with one confident channel per ray; labels are a subset of ids."""
    labels_present = torch.tensor(labels_present)
    lab = labels_present[torch.randint(0, len(labels_present), (n,), generator=gen)]
    logits = torch.randn(n, k, generator=gen) * 1.5
    perm = torch.randperm(k, generator=gen)                  # the network's channel order is arbitrary: that is what gets matched
    hot = perm[lab % k]
    logits[torch.arange(n), hot] += 4.0
    return torch.sigmoid(logits), lab.to(torch.int16).float()


def main():
    gen = torch.Generator().manual_seed(2024)
    save = {}
    cases = (("a", 512, 13, [0, 1, 2, 5, 7, 12]),            # fewer objects in the batch than channels: invalid_ce active
             ("b", 300, 59, list(range(0, 59, 3))),
             ("c", 257, 6, [0, 1, 2, 3, 4, 5]))              # every channel matched: invalid_ce = tensor([0])
    for tag, n, k, present in cases:
        pred, lab = case(gen, n, k, present)
        p1 = pred.clone().requires_grad_(True)
        ref = ref_ins_criterion(p1, lab, k)
        ref[0].sum().backward()
        p2 = pred.clone().requires_grad_(True)
        mine = O.ins_criterion(p2, lab, k)
        mine[0].sum().backward()
        for a, b in zip(ref, mine):
            assert torch.equal(a, b), "oracle != reference (ins_criterion %s)" % tag
        assert torch.equal(p1.grad, p2.grad), "oracle gradient != reference (%s)" % tag
        valid = torch.unique(lab)
        gt = torch.zeros(n, k)
        gt[:, :len(valid)] = torch.nn.functional.one_hot(lab.long())[..., valid.long()]
        ce, siou, rows, cols = ref_hungarian(pred, gt, len(valid), k)
        ce2, siou2, rows2, cols2 = O.hungarian(pred, gt, len(valid), k)
        assert torch.equal(ce, ce2) and torch.equal(siou, siou2) and list(rows) == list(rows2) and list(cols) == list(cols2)
        save.update({"pred_" + tag: pred.numpy(), "labels_" + tag: lab.numpy(), "k_" + tag: k,
                     "loss_" + tag: np.array([float(x.detach().sum()) for x in ref], dtype=np.float32),
                     "cost_ce_" + tag: ce.numpy(), "cost_siou_" + tag: siou.numpy(), "order_col_" + tag: np.asarray(cols, dtype=np.int64),
                     "grad_" + tag: p1.grad.numpy()})
        print("ins_criterion case %s: N %d K %d valid %d -> sum %.6f ce %.6f inv %.6f siou %.6f" %
              ((tag, n, k, len(valid)) + tuple(float(x.detach().sum()) for x in ref)))
    np.savez_compressed(os.path.join(OUT, "evaluator.npz"), **save)
    print("written", os.path.join(OUT, "evaluator.npz"), os.path.getsize(os.path.join(OUT, "evaluator.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
