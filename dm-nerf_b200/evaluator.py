"""networks/evaluator.py:19-74 on the native kernels: the Hungarian-matched instance loss of the training step.

    ins_criterion(pred_ins, gt_labels, ins_num) -> (ins_loss_sum, valid_ce, invalid_ce, valid_siou)     evaluator.py:19-37
    hungarian(pred_ins, gt_ins, valid_ins_num, ins_num) -> (cost_ce, cost_siou, order_row, order_col)   evaluator.py:41-74
    img2mse, mse2psnr, to8b                                                                             evaluator.py:11,14,15

The two [ins x ins] cost matrices come out of ONE pass over the rays (csrc/evaluator.cu: gt is one-hot, so the reference's
[ins x ins x N] broadcast collapses to per-row sums).  ins_criterion then runs the assignment ON THE DEVICE (scipy's
shortest-augmenting-path algorithm restated for one warp, same fp64 duals and tie rule: the training iteration has no
device->host hop left; DMNERF_INS_ASSIGN=host selects scipy on the host, the reference's own arrangement, with one hop per
call).  hungarian() returns the reference's numpy orders and therefore always uses scipy.  The matched loss and its gradient
w.r.t. the rendered instance map are evaluated on the device (dmnerf_hungarian_assign / dmnerf_ins_loss_backward[_dev]).
"""
import os
import numpy as np
import torch

from . import _lib
from .engine import get_context

img2mse = lambda x, y: torch.mean((x - y) ** 2)                                        # evaluator.py:11
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)                              # evaluator.py:14
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.tensor([10.], device=x.device))   # evaluator.py:15


def _costs(pred, gt_row):
    """pred [N,K] float32 contiguous, gt_row [N] int32 -> dict of device tensors (cost matrices + backward sums)."""
    n, k = pred.shape
    dev = pred.device
    e = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
    out = {"cost_ce": e(k, k), "cost_siou": e(k, k), "tp": e(k, k), "col_sum": e(k), "row_count": e(k)}
    ctx = get_context(dev)
    _lib.check(ctx.lib.dmnerf_hungarian_costs(_lib.ptr(pred), gt_row.data_ptr(), n, k, _lib.ptr(out["cost_ce"]),
                                              _lib.ptr(out["cost_siou"]), _lib.ptr(out["tp"]), _lib.ptr(out["col_sum"]),
                                              _lib.ptr(out["row_count"]), ctx.stream()), "dmnerf_hungarian_costs")
    return out


def _reorder(cost_matrix, valid_ins_num, ins_num):
    """evaluator.py:42-50 (host): assignment on the valid rows, unmatched prediction channels appended."""
    from scipy.optimize import linear_sum_assignment
    scores = cost_matrix[:valid_ins_num].detach().cpu().numpy()
    row_ind, col_ind = linear_sum_assignment(scores)
    if ins_num - valid_ins_num > 0:
        unmapped = np.array(list(set(range(ins_num)) - set(col_ind)))
        col_ind = np.concatenate([col_ind, unmapped])
    return row_ind, col_ind


def _rows_of_dense_gt(gt_ins):
    """Dense one-hot-or-zero gt_ins [N,K] -> row index per ray (-1 where the ray has no label)."""
    has = gt_ins.sum(-1) > 0
    return torch.where(has, gt_ins.argmax(-1), torch.full_like(has, -1, dtype=torch.int64)).to(torch.int32)


def hungarian(pred_ins, gt_ins, valid_ins_num, ins_num):
    """evaluator.py:41-74 for CUDA tensors (no gradient: use ins_criterion for the differentiable loss)."""
    if not pred_ins.is_cuda:
        raise RuntimeError("hungarian: expected CUDA tensors (no CPU fallback)")
    pred = pred_ins.detach().contiguous().float()
    gt_row = _rows_of_dense_gt(gt_ins.to(pred.device)).contiguous()
    c = _costs(pred, gt_row)
    order_row, order_col = _reorder(c["cost_ce"] + c["cost_siou"], valid_ins_num, ins_num)
    return c["cost_ce"], c["cost_siou"], order_row, order_col


class _MatchedLoss(torch.autograd.Function):
    """forward(pred_ins [N,K], gt_row [N] int32, valid_rows): gt_row[i] = cost-matrix row of ray i.  valid_rows = None: the rows
    are the label VALUES themselves (labels in [0, K)); which of them occur is read off the row populations that come back with
    the cost matrices -- ONE device->host hop per call.  valid_rows = n: rows 0..n-1 are the compacted labels (general path)."""

    @staticmethod
    def forward(fctx, pred_ins, gt_row, valid_rows):
        pred = pred_ins.detach().contiguous().float()
        n, k = pred.shape
        c = _costs(pred, gt_row)
        dev = pred.device
        if valid_rows is None:
            host = torch.cat([(c["cost_ce"] + c["cost_siou"]).reshape(-1), c["row_count"]]).cpu().numpy()   # the one sync
            scores_all, counts = host[:k * k].reshape(k, k), host[k * k:]
            if int(round(float(counts.sum()))) != n:
                raise _LabelsOutOfRange()
            rows_present = np.nonzero(counts > 0)[0]                                        # ascending = torch.unique order
            scores = scores_all[rows_present]
        else:
            rows_present = np.arange(int(valid_rows))
            scores = (c["cost_ce"] + c["cost_siou"])[:int(valid_rows)].cpu().numpy()
        n_valid = len(rows_present)
        from scipy.optimize import linear_sum_assignment
        row_ind, col_ind = linear_sum_assignment(scores)                                    # evaluator.py:45-47
        unmatched = np.array(sorted(set(range(k)) - set(col_ind.tolist())), dtype=np.int64)
        rows = torch.as_tensor(rows_present[row_ind], device=dev, dtype=torch.int64)
        cols = torch.as_tensor(col_ind, device=dev, dtype=torch.int64)
        valid_ce = c["cost_ce"][rows, cols].mean()                                          # evaluator.py:28
        valid_siou = c["cost_siou"][rows, cols].mean()                                      # evaluator.py:34
        row_of_col = torch.full((k,), -1, device=dev, dtype=torch.int32)
        row_of_col[cols] = rows.to(torch.int32)
        if len(unmatched):                                                                  # evaluator.py:30-33
            un = torch.as_tensor(unmatched, device=dev, dtype=torch.int64)
            invalid_ce = c["col_sum"][un].sum() / float(n * len(un))
        else:
            invalid_ce = torch.zeros((), device=dev)
        fctx.save_for_backward(pred, gt_row, row_of_col, c["tp"], c["col_sum"], c["row_count"])
        fctx.n_valid = int(n_valid)
        fctx.in_shape = pred_ins.shape
        n_valid_t = torch.tensor(int(n_valid))                                               # host-side by-product, not differentiable
        fctx.mark_non_differentiable(n_valid_t)
        return valid_ce, invalid_ce, valid_siou, n_valid_t

    @staticmethod
    def backward(fctx, g_ce, g_inv, g_siou, _g_n=None):
        pred, gt_row, row_of_col, tp, col_sum, row_count = fctx.saved_tensors
        n, k = pred.shape
        zero = pred.new_zeros(())
        g3 = torch.stack([(g if g is not None else zero).reshape(()).float() for g in (g_ce, g_inv, g_siou)]).contiguous()
        d_pred = torch.empty_like(pred)
        ctx = get_context(pred.device)
        _lib.check(ctx.lib.dmnerf_ins_loss_backward(_lib.ptr(pred), gt_row.data_ptr(), n, k, row_of_col.data_ptr(), fctx.n_valid,
                                                    _lib.ptr(tp), _lib.ptr(col_sum), _lib.ptr(row_count), _lib.ptr(g3),
                                                    _lib.ptr(d_pred), ctx.stream()), "dmnerf_ins_loss_backward")
        return d_pred.reshape(fctx.in_shape), None, None


class _LabelsOutOfRange(Exception):
    pass


class _MatchedLossDevice(torch.autograd.Function):
    """forward(pred_ins [N,K], labels [N] int32): ranks of the labels, cost matrices, assignment and the three loss terms in four
    launches, nothing read back.  Labels outside [0, 65536) or more distinct labels than K: NaN losses, zero gradient, and the
    next ins_criterion call raises (error word in mapped host memory)."""

    @staticmethod
    def forward(fctx, pred_ins, labels):
        pred = pred_ins.detach().contiguous().float()
        n, k = pred.shape
        dev = pred.device
        ctx = get_context(dev)
        gt_row = torch.empty(n, device=dev, dtype=torch.int32)
        n_valid = torch.empty(1, device=dev, dtype=torch.int32)
        _lib.check(ctx.lib.dmnerf_ins_label_rows(labels.data_ptr(), n, k, gt_row.data_ptr(), n_valid.data_ptr(), ctx.stream()),
                   "dmnerf_ins_label_rows")
        c = _costs(pred, gt_row)
        row_of_col = torch.empty(k, device=dev, dtype=torch.int32)
        losses = torch.empty(3, device=dev, dtype=torch.float32)
        _lib.check(ctx.lib.dmnerf_hungarian_assign(_lib.ptr(c["cost_ce"]), _lib.ptr(c["cost_siou"]), _lib.ptr(c["col_sum"]),
                                                   n_valid.data_ptr(), n, k, row_of_col.data_ptr(), _lib.ptr(losses), ctx.stream()),
                   "dmnerf_hungarian_assign")
        fctx.save_for_backward(pred, gt_row, row_of_col, n_valid, c["tp"], c["col_sum"], c["row_count"])
        fctx.in_shape = pred_ins.shape
        fctx.mark_non_differentiable(n_valid, row_of_col)
        valid_ce, invalid_ce, valid_siou = losses[0].clone(), losses[1].clone(), losses[2].clone()
        return valid_ce, invalid_ce, valid_siou, n_valid, row_of_col

    @staticmethod
    def backward(fctx, g_ce, g_inv, g_siou, _g_n=None, _g_r=None):
        pred, gt_row, row_of_col, n_valid, tp, col_sum, row_count = fctx.saved_tensors
        n, k = pred.shape
        zero = pred.new_zeros(())
        g3 = torch.stack([(g if g is not None else zero).reshape(()).float() for g in (g_ce, g_inv, g_siou)]).contiguous()
        d_pred = torch.empty_like(pred)
        ctx = get_context(pred.device)
        _lib.check(ctx.lib.dmnerf_ins_loss_backward_dev(_lib.ptr(pred), gt_row.data_ptr(), n, k, row_of_col.data_ptr(),
                                                        n_valid.data_ptr(), _lib.ptr(tp), _lib.ptr(col_sum), _lib.ptr(row_count),
                                                        _lib.ptr(g3), _lib.ptr(d_pred), ctx.stream()), "dmnerf_ins_loss_backward_dev")
        return d_pred.reshape(fctx.in_shape), None


_STATUS_TEXT = {701: "a label outside [0, 65536)", 702: "more distinct labels than ins_num (or an empty batch)"}


def ins_assignment(pred_ins, gt_labels, ins_num):
    """Device-side matching only: (row_of_col [ins_num] int32: rank of the matched label per prediction channel or -1,
    n_valid [1] int32), both on the device.  Same kernels as ins_criterion; for tests and diagnostics."""
    labels = gt_labels.to(pred_ins.device).reshape(-1).to(torch.int32).contiguous()
    out = _MatchedLossDevice.apply(pred_ins, labels)
    return out[4], out[3]


def ins_criterion(pred_ins, gt_labels, ins_num):
    """evaluator.py:19-37.  pred_ins [N, ins_num] (rendered instance probabilities, CUDA), gt_labels [N] (object ids)."""
    if not pred_ins.is_cuda:
        raise RuntimeError("ins_criterion: expected CUDA tensors (no CPU fallback)")
    if pred_ins.dim() != 2 or pred_ins.shape[1] != ins_num or gt_labels.shape[0] != pred_ins.shape[0]:
        raise RuntimeError("ins_criterion: pred_ins %s / gt_labels %s / ins_num %d are inconsistent"
                           % (tuple(pred_ins.shape), tuple(gt_labels.shape), ins_num))
    labels = gt_labels.to(pred_ins.device).reshape(-1)
    if os.environ.get("DMNERF_INS_ASSIGN", "device") != "host":
        code = get_context(pred_ins.device).lib.dmnerf_ins_status_take()
        if code:
            raise RuntimeError("ins_criterion: an earlier call was given %s (code %d); its loss was NaN and its gradient zero. "
                               "DMNERF_INS_ASSIGN=host handles arbitrary integer labels (one synchronisation per call)."
                               % (_STATUS_TEXT.get(code, "labels it cannot rank"), code))
        valid_ce, invalid_ce, valid_siou, _, _ = _MatchedLossDevice.apply(pred_ins, labels.to(torch.int32).contiguous())
        # evaluator.py:33 returns tensor([0]) (shape [1]) when every channel is matched; here invalid_ce is a 0-dim zero in that
        # case (the number of distinct labels is not known on the host)
        return valid_ce + invalid_ce + valid_siou, valid_ce, invalid_ce, valid_siou             # evaluator.py:36
    try:
        # object ids in [0, ins_num) (every dataset of the reference): the id IS the cost-matrix row; which ids occur comes back
        # with the matrices, so the call makes a single device->host hop (no torch.unique synchronisation)
        valid_ce, invalid_ce, valid_siou, n_valid_t = _MatchedLoss.apply(pred_ins, labels.to(torch.int32).contiguous(), None)
    except _LabelsOutOfRange:
        valid = torch.unique(labels)                                                        # evaluator.py:21 (sorted)
        n_valid = int(valid.numel())
        if n_valid > ins_num:
            raise RuntimeError("ins_criterion: %d distinct labels for ins_num %d" % (n_valid, ins_num))
        gt_row = torch.searchsorted(valid, labels).to(torch.int32).contiguous()             # column of the one-hot, :25
        valid_ce, invalid_ce, valid_siou, n_valid_t = _MatchedLoss.apply(pred_ins, gt_row, n_valid)
    if int(n_valid_t) == ins_num:
        invalid_ce = torch.tensor([0], device=pred_ins.device)                              # evaluator.py:33
    ins_loss_sum = valid_ce + invalid_ce + valid_siou                                       # evaluator.py:36
    return ins_loss_sum, valid_ce, invalid_ce, valid_siou
