"""CPU restatement of scipy.optimize.linear_sum_assignment -- TEST INFRASTRUCTURE ONLY.

The reference calls scipy for the Hungarian assignment of its instance loss (networks/evaluator.py:43-45, scipy pinned to
1.7.1 in the reference's requirements.txt:103); scipy is a third-party dependency, not part of /root/reference.  Its solver is
the rectangular shortest-augmenting-path algorithm of D. F. Crouse, "On implementing 2D rectangular assignment algorithms",
IEEE Trans. Aerospace and Electronic Systems 52(4), 2016 (scipy/optimize/rectangular_lsap/rectangular_lsap.cpp, unchanged in
its arithmetic and tie rule since scipy 1.6).  The product runs the same algorithm on the device
(dm-nerf_b200/csrc/evaluator.cu::hungarian_assign_kernel); this file restates it twice:

  lsap_sequential(cost)     the published algorithm column by column, as scipy's C++ does it: fp64 duals, remaining columns kept
                            in a vector filled in reverse order and compacted by swap-with-last, and among equal path costs a
                            later column replaces the current choice only if it is unassigned;
  lsap_lane_parallel(cost)  the form the kernel uses: all remaining columns of a scan are evaluated independently and ONE
                            arg-min picks the column -- equal costs are ordered by a unique integer score per scan position
                            (unassigned: R + position, assigned: R - 1 - position; larger wins), which selects the LAST
                            unassigned minimum if there is one and the FIRST minimum otherwise: exactly the column the
                            sequential scan ends up with.

Pin: tests/test_oracle.py checks both against the installed scipy on random, small-integer (tie-heavy), constant and
duplicate-column matrices -- the same columns, not just the same total cost; tests/test_gpu_train.py checks the kernel against
scipy on the same families.  Only tests/ may import this module.
"""
import numpy as np

INF = float("inf")


def _solve(cost, pick):
    cost = np.asarray(cost, dtype=np.float64)
    nr, nc = cost.shape
    if nr > nc:
        raise ValueError("lsap: more rows than columns (scipy transposes; the instance loss never has this shape)")
    u, v = np.zeros(nr), np.zeros(nc)
    path = -np.ones(nc, dtype=np.int64)
    col4row = -np.ones(nr, dtype=np.int64)
    row4col = -np.ones(nc, dtype=np.int64)
    for cur in range(nr):
        remaining = [nc - it - 1 for it in range(nc)]          # reverse fill: a constant matrix gives the identity
        n_rem = nc
        in_sr = np.zeros(nr, dtype=bool)
        in_sc = np.zeros(nc, dtype=bool)
        sp = np.full(nc, INF)
        min_val, i, sink = 0.0, cur, -1
        while sink == -1:
            in_sr[i] = True
            for it in range(n_rem):                               # relax every remaining column through row i
                j = remaining[it]
                r = ((min_val + cost[i, j]) - u[i]) - v[j]      # scipy's left-to-right evaluation
                if r < sp[j]:
                    path[j] = i
                    sp[j] = r
            index = pick(remaining, n_rem, sp, row4col)
            min_val = sp[remaining[index]]
            if min_val == INF:
                raise ValueError("cost matrix is infeasible")
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            in_sc[j] = True
            n_rem -= 1
            remaining[index] = remaining[n_rem]
        u[cur] += min_val
        for i2 in range(nr):
            if in_sr[i2] and i2 != cur:
                u[i2] += min_val - sp[col4row[i2]]
        for j in range(nc):
            if in_sc[j]:
                v[j] -= min_val - sp[j]
        j = sink
        while True:                                               # augment along the alternating path
            i2 = path[j]
            row4col[j] = i2
            col4row[i2], j = j, col4row[i2]
            if i2 == cur:
                break
    return np.arange(nr), col4row.copy()


def _pick_sequential(remaining, n_rem, sp, row4col):
    index, lowest = -1, INF
    for it in range(n_rem):
        j = remaining[it]
        if sp[j] < lowest or (sp[j] == lowest and row4col[j] == -1):
            lowest, index = sp[j], it
    return index if index >= 0 else 0


def _pick_lane_parallel(remaining, n_rem, sp, row4col):
    best = None
    for it in range(n_rem):                                       # any evaluation order gives the same winner
        j = remaining[it]
        score = n_rem + it if row4col[j] == -1 else n_rem - 1 - it
        key = (sp[j], -score)
        if best is None or key < best[0]:
            best = (key, it)
    return best[1]


def lsap_sequential(cost):
    """(row_ind, col_ind) of scipy.optimize.linear_sum_assignment(cost) for rows <= columns."""
    return _solve(cost, _pick_sequential)


def lsap_lane_parallel(cost):
    """The same assignment from the order-free arg-min the device kernel uses."""
    return _solve(cost, _pick_lane_parallel)


def tie_heavy_cases(n_cases, max_n, seed):
    """Score matrices of the four families the tests use: uniform random, small integers, constant, duplicated columns."""
    rng = np.random.default_rng(seed)
    for trial in range(n_cases):
        nc = int(rng.integers(1, max_n + 1))
        nr = int(rng.integers(1, nc + 1)) if trial % 3 else nc
        kind = trial % 4
        if kind == 0:
            c = rng.random((nr, nc))
        elif kind == 1:
            c = rng.integers(0, 3, (nr, nc)).astype(np.float64)
        elif kind == 2:
            c = np.full((nr, nc), float(rng.integers(0, 2)))
        else:
            c = rng.random((nr, nc))
            c[:, rng.integers(0, nc, max(1, nc // 2))] = c[:, [0]]
        yield c.astype(np.float32)
