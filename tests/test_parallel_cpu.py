"""CPU, world_size 2 over gloo: the ray sharding + single all-gather assembly of dm-nerf_b200/parallel.py, with the
oracle standing in for the CUDA renderer (host-side logic only)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dmnerf_b200 import synth
    from dmnerf_b200.parallel import render_frame_sharded, shard_range
    from oracle import dmnerf_oracle as O
    torch.set_num_threads(2)
    wl = synth.workload("dmsr_study")
    sel = np.linspace(0, 307199, 301).astype(np.int64)            # ragged: 301 rays over 2 ranks, shards of 256/45
    ro, rd = torch.from_numpy(wl["rays_o"][sel]), torch.from_numpy(wl["rays_d"][sel])
    pc, pf = O.to_torch(synth.make_weights(101, 13)), O.to_torch(synth.make_weights(202, 13))

    def oracle_render(o, d, mc, mf, z, N_importance=128, want_raw=False, want_coarse=False):
        zz = O.z_val_sample(o.shape[0], wl["near"], wl["far"], 64)
        r = O.render(o, d, pc, pf, zz, n_importance=N_importance)
        return {k: r[k] for k in ("rgb_fine", "depth_fine", "acc_fine", "ins_fine")}

    with torch.no_grad():
        img = render_frame_sharded(ro, rd, None, None, None, render_fn=oracle_render)
        if rank == 0:
            ref = oracle_render(ro, rd, None, None, None)
            ok = all(torch.allclose(img[k], ref[k], rtol=1e-5, atol=1e-6) for k in ref)
            shapes = {k: tuple(v.shape) for k, v in img.items()}
            ret.put((ok, shapes, shard_range(301, 2, 0, 128), shard_range(301, 2, 1, 128)))
    dist.barrier()
    dist.destroy_process_group()


def _camera_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dmnerf_b200.parallel import render_trajectory_sharded
    H, W, ins = 9, 23, 5                                       # 207 pixels: shards of 128 / 79

    def fake_frame(H, W, K, c2w, near, far, mc, mf, N_samples=64, N_importance=128, pixel_range=None):
        lo, cnt = pixel_range
        idx = torch.arange(lo, lo + cnt, dtype=torch.float32) + 1000.0 * float(c2w)
        return {"rgb": torch.stack([idx, idx + 0.25, idx + 0.5], -1), "ins": idx[:, None] * torch.arange(1, ins + 1),
                "depth": -idx, "acc": idx * 2}

    imgs = list(render_trajectory_sharded([1.0, 2.0], H, W, None, 0.0, 1.0, None, None, frame_fn=fake_frame))
    if rank == 1:                                              # every rank holds the full image
        ok = True
        for pose, img in zip((1.0, 2.0), imgs):
            idx = torch.arange(H * W, dtype=torch.float32) + 1000.0 * pose
            ok &= torch.equal(img["rgb"].reshape(-1, 3)[:, 0], idx) and torch.equal(img["depth"].reshape(-1), -idx)
            ok &= torch.equal(img["ins"].reshape(-1, ins)[:, 2], idx * 3) and torch.equal(img["acc"].reshape(-1), idx * 2)
            ok &= img["rgb"].shape == (H, W, 3) and img["ins"].shape == (H, W, ins)
        ret.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_trajectory_render_sharded_by_pixel_range_gloo_world2():
    """BASELINE config 5 host logic: every pose is split by pixel range over the ranks (frame driver per rank) and assembled
    with one all-gather; world 2 over gloo with a synthetic frame function."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_camera_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok


@pytest.mark.timeout(600)
def test_sharded_render_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, shapes, s0, s1 = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok
    assert shapes == {"rgb_fine": (301, 3), "depth_fine": (301,), "acc_fine": (301,), "ins_fine": (301, 13)}
    assert s0 == (0, 256, 256) and s1 == (256, 301, 256)


def test_shard_range_covers_everything():
    from dmnerf_b200.parallel import shard_range
    for n in (0, 1, 127, 128, 4096, 307200, 307201):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi, per = shard_range(n, world, r, 128)
                assert per % 128 == 0 and hi - lo <= per
                got += list(range(lo, hi))
            assert got == list(range(n))
