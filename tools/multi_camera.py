"""2+ GPU check of the sharded camera render (BASELINE config 5 path): every rank renders its pixel range through the frame
driver, one NCCL all-gather assembles the image, and rank 0 compares with the whole frame rendered locally.
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_camera.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmnerf_b200 import synth                                   # noqa: E402
from dmnerf_b200.testing import make_models                     # noqa: E402
from dmnerf_b200.render import render_frame                     # noqa: E402
from dmnerf_b200.parallel import render_trajectory_sharded     # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
wl = synth.workload("replica_office2")
ins_num = int(wl["ins_num"])
nc, nf, _, _ = make_models(5, 6, ins_num, dev)
H, W = 120, 160
K = np.array(wl["K"], dtype=np.float32).copy()
K[0, 2], K[1, 2] = W / 2 - 0.5, H / 2 - 0.5
K[0, 0] = K[1, 1] = 80.0
poses = [synth.workload("replica_office2", frame=f)["c2w"] for f in (0, 7, 450)]
near, far = float(wl["near"]), float(wl["far"])
with torch.no_grad():
    t0 = time.perf_counter()
    imgs = list(render_trajectory_sharded(poses, H, W, K, near, far, nc, nf, device=dev))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = True
    if rank == 0:
        for c2w, img in zip(poses, imgs):
            ref = render_frame(H, W, K, c2w, near, far, nc, nf, device=dev)
            for k in ("rgb", "ins", "depth", "acc"):
                ok &= torch.equal(img[k].cpu(), ref[k])
        print("multi_camera: world %d, %d poses of %dx%d (ins_num %d) in %.3f s, identical to the single-GPU frames: %s"
              % (world, len(poses), W, H, ins_num, dt, ok))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
