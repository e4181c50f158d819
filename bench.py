#!/usr/bin/env python
"""Benchmark of the DM-NeRF render hot path (BASELINE.json metric: rays/sec, 64 coarse + 128 fine samples).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one full 640x480 synthetic frame (307 200 rays) per GPU through the complete pipeline
(coarse net -> composite -> importance sampling -> fine net -> composite, object head included), i.e.
BASELINE.json configs[1] (DM-SR 'study').  Under torchrun (N > 1) every rank renders its own frame of the
synthetic trajectory (rays shard with no data-path collective; weak scaling) and the rendered images are
all-gathered once per step over NCCL, inside the timed region.

Timed quantities (ours):
  value      rays/s, device-timed (CUDA events on the launch stream), inputs already resident in HBM
  e2e        rays/s through the C-ABI host entry point (dmnerf_render_forward_host): pinned HOST rays in,
             HOST rgb/depth/acc/ins out, H2D + D2H inside the timed region
  roofline   the dominant kernel (fine-network MLP): algorithmic FLOPs / its CUDA-event duration vs the measured
             tensor peak in MEASURED_PEAKS.json
  cpu_baseline  the oracle port (torch CPU restatement of the reference) on this box's host cores, bounded sample

--impl reference times that same CPU path alone (the reference is pure Python/torch and cannot travel to the GPU
box; the oracle is its bit-for-bit restatement, see oracle/dmnerf_oracle.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

N_COARSE, N_IMPORTANCE = 64, 128
_CPU_THREADS = None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dmsr_study")
    ap.add_argument("--mlp", default="auto", choices=["auto", "simt", "umma"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_rate(workload, seconds, chunk=1024):
    """rays/s of the oracle port on the host cores: full dm_nerf pipeline, 1024-ray chunks (BASELINE configs[0]
    chunk size), repeated until `seconds` of CPU work have elapsed (at least 2 chunks after 1 warm-up)."""
    from oracle import dmnerf_oracle as O
    from dmnerf_b200 import synth
    wl = synth.workload(workload)
    cores = os.cpu_count() or 1
    pc = O.to_torch(synth.make_weights(101, wl["ins_num"]))
    pf = O.to_torch(synth.make_weights(202, wl["ins_num"]))
    ro, rd = torch.from_numpy(wl["rays_o"]), torch.from_numpy(wl["rays_d"])
    z = O.z_val_sample(chunk, wl["near"], wl["far"], N_COARSE)
    n_total = ro.shape[0]
    with torch.no_grad():
        # the reference would run with torch's default (= all cores); on many-core hosts that oversubscribes these
        # small GEMMs badly, so give the CPU arm its best thread count (quick calibration on 256 rays)
        global _CPU_THREADS
        if _CPU_THREADS is None:
            best = None
            for th in sorted({cores, min(cores, 64), min(cores, 32), min(cores, 16), min(cores, 8)}, reverse=True):
                torch.set_num_threads(th)
                O.render(ro[:256], rd[:256], pc, pf, z[:256], perturb=0.0, n_importance=N_IMPORTANCE)
                t0 = time.perf_counter()
                O.render(ro[256:512], rd[256:512], pc, pf, z[:256], perturb=0.0, n_importance=N_IMPORTANCE)
                dt = time.perf_counter() - t0
                if best is None or dt < best[0]:
                    best = (dt, th)
            _CPU_THREADS = best[1]
        torch.set_num_threads(_CPU_THREADS)
        O.render(ro[:chunk], rd[:chunk], pc, pf, z, perturb=0.0, n_importance=N_IMPORTANCE)     # warm-up
        done, t0 = 0, time.perf_counter()
        pos = chunk
        while True:
            O.render(ro[pos:pos + chunk], rd[pos:pos + chunk], pc, pf, z, perturb=0.0, n_importance=N_IMPORTANCE)
            done += chunk
            pos = (pos + chunk) % (n_total - chunk)
            el = time.perf_counter() - t0
            if el >= seconds and done >= 2 * chunk:
                break
    return done / el, cores, torch.get_num_threads(), done, el


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dmnerf_b200 import synth
    wl = synth.workload(args.workload)
    rates = []
    per_step = max(2.0, min(20.0, 100.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        r, cores, threads, rays, el = cpu_reference_rate(args.workload, per_step)
        if i >= args.warmup:
            rates.append((r, rays, el))
    tot_rays = sum(x[1] for x in rates)
    tot_t = sum(x[2] for x in rates)
    value = tot_rays / tot_t
    line = {
        "impl": "reference", "metric": "rays/sec (64c+128f samples)", "value": value, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DM-SR 'study' 640x480 full render, 64+128 hierarchical, coarse+fine+object head "
                               "(ins_num=%d), one frame (%d rays) per GPU per step" % (wl["ins_num"], wl["rays_o"].shape[0]),
                   "name": args.workload, "rays_per_step_per_gpu": int(wl["rays_o"].shape[0]),
                   "note": "CPU arm: each step renders a bounded sample of that frame (1024-ray chunks, as the reference's "
                           "tester.py chunk loop would) and reports rays/s"},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": "%d rays per step in 1024-ray chunks, torch CPU fp32, %d threads" % (rates[-1][1], threads)},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        reasons = []
        for i, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(r[i].lower().startswith("active") for r in rows):
                reasons.append(name)
        try:
            pw = max(float(r[2]) for r in rows)
        except ValueError:
            pw = None
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": pw, "samples": len(rows),
                "reasons": reasons}


# --------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import ctypes as C
    from dmnerf_b200 import synth, _lib
    from dmnerf_b200.engine import get_context
    from dmnerf_b200.testing import make_models

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (ours) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    impl = {"auto": _lib.IMPL_AUTO, "simt": _lib.IMPL_SIMT, "umma": _lib.IMPL_UMMA}[args.mlp]

    wl = synth.workload(args.workload, frame=rank)                  # each rank: its own frame of the trajectory
    ins_num = wl["ins_num"]
    n_rays = wl["rays_o"].shape[0]
    nc, nf, _, _ = make_models(101, 202, ins_num, dev)
    ctx = get_context(dev)
    ctx.bind(0, nc); ctx.bind(1, nf)
    from dmnerf_b200.render import render_rays
    from dmnerf_b200.parallel import gather_image
    ro, rd = torch.from_numpy(wl["rays_o"]).to(dev), torch.from_numpy(wl["rays_d"]).to(dev)
    z = (torch.linspace(0, 1, N_COARSE) * (wl["far"] - wl["near"]) + wl["near"]).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step_device():
        out = render_rays(ro, rd, nc, nf, z, want_raw=False, want_coarse=False, want_samples=False, impl=impl)
        if world > 1:
            return gather_image(out, world)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up
    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            step_device()
        barrier()
        # ---- timed: K steps, per-step CUDA events (L2 flushed, untimed, between steps)
        launches0 = _lib.launch_count()
        lib = ctx.lib
        _lib.check(lib.dmnerf_profile_enable(ctx.handle, 1), "profile_enable")
        sampler = ClockSampler(local)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        stage_ms = np.zeros(6)
        barrier()
        t_wall0 = time.perf_counter()
        for s in range(args.steps):
            flush.fill_(s & 0xFF)
            evs[s][0].record()
            step_device()
            evs[s][1].record()
            ms = (C.c_float * 6)()
            _lib.check(lib.dmnerf_profile_read(ctx.handle, ms, 6), "profile_read")
            stage_ms += np.array(list(ms))
        barrier()
        t_wall1 = time.perf_counter()
        clocks = sampler.stop(t_wall0, t_wall1)
        _lib.check(lib.dmnerf_profile_enable(ctx.handle, 0), "profile_enable")
        launches = _lib.launch_count() - launches0
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())

        # ---- end to end through the C ABI with HOST buffers (pinned), H2D + D2H inside the timed region
        h = {k: torch.empty(shape, dtype=torch.float32).pin_memory() for k, shape in
             (("rgb_fine", (n_rays, 3)), ("depth_fine", (n_rays,)), ("acc_fine", (n_rays,)), ("ins_fine", (n_rays, ins_num)))}
        hro, hrd = torch.from_numpy(wl["rays_o"]).pin_memory(), torch.from_numpy(wl["rays_d"]).pin_memory()
        hz = z.cpu().pin_memory()
        io = _lib.RenderIO()
        io.rays_o, io.rays_d, io.z_coarse, io.z_row_stride = _lib.ptr(hro), _lib.ptr(hrd), _lib.ptr(hz), 0
        for k, v in h.items():
            setattr(io, k, _lib.ptr(v))
        h2d = hro.numel() * 4 + hrd.numel() * 4 + hz.numel() * 4
        d2h = sum(v.numel() * 4 for v in h.values())

        def step_host():
            _lib.check(lib.dmnerf_render_forward_host(ctx.handle, io, n_rays, N_COARSE, N_IMPORTANCE, 0, impl, ctx.stream()),
                       "dmnerf_render_forward_host")

        for _ in range(2):
            step_host()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            step_host()
        e1.record()
        barrier()
        tw1 = time.perf_counter()
        e2e_ms = max(e0.elapsed_time(e1), 1e3 * (tw1 - tw0))        # host-synchronous call: wall clock is the honest one
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())

    # ---- BASELINE config 4 (informational): training step, 1024 random rays, forward + backward through the native path
    train = None
    if world == 1:
        import types
        from dmnerf_b200.render import dm_nerf
        from dmnerf_b200.embedder import get_embedder
        sel = torch.from_numpy(np.random.Generator(np.random.PCG64(0)).choice(n_rays, 1024, replace=False)).to(dev)
        rays = torch.stack([ro[sel], rd[sel]], 0)
        targs = types.SimpleNamespace(perturb=1.0, N_importance=N_IMPORTANCE, is_train=True, N_ins=None)
        pe, ve = get_embedder(10)[0], get_embedder(4)[0]
        zc = z[None].expand(1024, N_COARSE)
        tgt = torch.rand(1024, 3, device=dev)
        nc.train(); nf.train()
        opt = torch.optim.Adam(list(nc.parameters()) + list(nf.parameters()), lr=5e-4)

        def train_step():
            out = dm_nerf(rays, pe, ve, nc, nf, zc, targs)
            loss = ((out["rgb_fine"] - tgt) ** 2).mean() + ((out["rgb_coarse"] - tgt) ** 2).mean() + out["ins_fine"].mean()
            opt.zero_grad()
            loss.backward()
            opt.step()

        for _ in range(2):
            train_step()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(5):
            train_step()
        torch.cuda.synchronize(dev)
        tms = 1e3 * (time.perf_counter() - t0) / 5
        train = {"rays_per_step": 1024, "ms_per_step": tms, "rays_per_s": 1024 / (tms * 1e-3),
                 "what": "dm_nerf(perturb=1) forward (tcgen05 kernel, activations saved from the epilogue) + backward (composite "
                         "reverse scan, split-bf16 tcgen05 GEMM kernels for dX / dW) + Adam step; wall clock"}
        # the emptiness penalizer on the fine network's per-sample outputs (forward + backward), HBM-streaming
        try:
            from dmnerf_b200.penalizer import ins_penalizer
            pargs = types.SimpleNamespace(tolerance=0.05, deta_w=0.05)
            praw = torch.randn(1024, N_COARSE + N_IMPORTANCE, 4 + ins_num + 1, device=dev, requires_grad=True)
            pz = (torch.rand(1024, N_COARSE + N_IMPORTANCE, device=dev).sort(-1).values * 11 + 4)
            pdep = pz[:, 100].clone()
            for _ in range(3):
                praw.grad = None
                ins_penalizer(praw, pz, pdep, rays[1], pargs).sum().backward()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(20):
                praw.grad = None
                ins_penalizer(praw, pz, pdep, rays[1], pargs).sum().backward()
            ev1.record()
            torch.cuda.synchronize(dev)
            pus = 1e3 * ev0.elapsed_time(ev1) / 20
            k1 = ins_num + 1
            pbytes = 1024 * (N_COARSE + N_IMPORTANCE) * (4 * (2 * k1 + k1) + 3 * 4)
            train["penalizer"] = {"us_per_fwd_bwd": pus, "algorithmic_GBps": pbytes / (pus * 1e-6) / 1e9,
                                  "what": "ins_penalizer on raw_fine [1024,192,%d]: count + loss kernels, gradient kernel (+ torch "
                                          "zero-fill of d_raw); algorithmic bytes = raw logits read twice + gradient written once + "
                                          "z_vals three times" % (4 + k1)}
        except Exception as exc:                       # informational block: never fail the bench line because of it
            train["penalizer"] = {"error": str(exc)}
        nc.eval(); nf.eval()

    total_rays = n_rays * world * args.steps
    value = total_rays / (dev_ms * 1e-3)
    e2e_value = total_rays / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fine network: F = 192 samples per ray)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peaks = json.load(open(peaks_path))
        peak_tf, peak_src = float(peaks["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
    else:
        peak_tf, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained)"
    fused = stage_ms[4] < 0.05 * stage_ms[0]                    # single-kernel pipeline: only stage 0 carries time
    if fused:
        fine_ms = stage_ms[0] / args.steps
        fine_flops = synth.flops_per_ray(ins_num) * n_rays
        kname = "fused render kernel (coarse 64 + fine 192 network evaluations per ray, composite, sampling)"
    else:
        fine_ms = stage_ms[4] / args.steps
        fine_flops = 2.0 * (N_COARSE + N_IMPORTANCE) * synth.macs_per_sample(ins_num) * n_rays
        kname = "fine-network MLP kernel (192 samples/ray)"
    achieved_tf = fine_flops / (fine_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch") if fused else None
    roofline = {"bound": "tensor", "kernel": kname, "achieved": achieved_tf,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": traffic,
                "peak_source": peak_src,
                "note": "algorithmic FLOPs (2*693504 MACs per network evaluation, the reference's layer shapes); the kernel issues 3 bf16 "
                        "tensor passes (hi/lo operand split for fp32 parity) over 563200 padded+folded MACs per sample, so "
                        "frac cannot exceed 693504/(3*563200) = 0.410 of the bf16 peak",
                "stage_ms_per_step": {k: float(v / args.steps) for k, v in
                                      zip(("coarse_z_or_fused_kernel", "coarse_mlp", "coarse_composite", "hier_sample", "fine_mlp",
                                           "fine_composite"), stage_ms)}}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        r, cores, threads, rays, el = cpu_reference_rate(args.workload, args.cpu_seconds)
        cpu_baseline = {"value": r, "unit": "rays/s", "cores": cores, "kind": "port",
                        "sample": "%d rays (1024-ray chunks, full 64+128 pipeline) in %.1f s, torch CPU fp32, %d threads"
                                  % (rays, el, threads)}

    line = {
        "metric": "rays/sec (64c+128f samples)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (bf16x3 split operands on tcgen05, fp32 accumulate)" if args.mlp != "simt" else "f32",
        "data": "synthetic",
        "config": {"workload": "DM-SR 'study' 640x480 full render, 64+128 hierarchical, coarse+fine+object head "
                               "(ins_num=%d), one frame (%d rays) per GPU per step" % (ins_num, n_rays),
                   "name": args.workload, "rays_per_step_per_gpu": n_rays, "l2": "256 MiB write between timed steps",
                   "mlp_impl": args.mlp, "parallelism": "rays sharded by frame, dp%d, one NCCL all-gather of the image per step" % world},
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps, "api": "dmnerf_render_forward_host (C ABI, pinned host buffers)"},
        "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
    }
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    if train:
        line["train_step"] = train
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
