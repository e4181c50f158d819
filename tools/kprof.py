"""In-kernel cycle profile of the tcgen05 MLP / fused render kernel (diagnostics build, -DDMN_KPROF).

  python tools/kprof.py --build     (here: compiles tools/bin/libdmnerf_kprof.so)
  python tools/kprof.py [--fused]   (on the GPU box)

Prints, for a few CTAs, where the MMA-issuing warp and one epilogue thread spent their cycles.
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tools", "bin", "libdmnerf_kprof.so")
NAMES = ["mma:wait acc drained", "mma:wait epi(g-1) chunk0", "mma:wait epi(g-1) chunk1", "mma:wait inputs", "mma:wait W_hi stage",
         "mma:wait W_lo stage", "mma:role total", "mma:wait (head steps)", "epi:wait acc_full", "epi:wait a_free",
         "epi:prologue", "epi:role total", "epi(x18): wait::st+fence+arrive", "epi(odd x9): split + st issue", "epi(odd x9): acc_full->ld done", "epi(odd x9): bias/relu"]


def build():
    global LIB
    extra = [x for x in sys.argv if x.startswith("-D")]
    if extra:
        LIB = LIB.replace(".so", "_" + "_".join(x[2:].lower() for x in extra) + ".so")
    sys.path.insert(0, ROOT)
    from dmnerf_b200 import build as b
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objs = []
    for src in b.SOURCES:
        obj = os.path.join(ROOT, "tools", "bin", "kprof_" + src.replace(".cu", ".o"))
        subprocess.check_call([b.nvcc()] + b.NVCC_FLAGS + extra + ["-DDMN_KPROF", "-I", os.path.join(ROOT, "include"), "-c",
                                                           os.path.join(b.CSRC, src), "-o", obj])
        objs.append(obj)
    subprocess.check_call([b.nvcc()] + b.NVCC_FLAGS[:2] + ["-shared", "-o", LIB] + objs + ["-lcudart"])
    print(LIB)


def main():
    if "--build" in sys.argv:
        return build()
    os.environ["DMNERF_LIB_PATH"] = os.environ.get("KPROF_LIB", LIB)
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from dmnerf_b200 import synth, _lib
    from dmnerf_b200.engine import get_context
    from dmnerf_b200.testing import model_from_weights, make_models
    from dmnerf_b200.autograd import mlp_forward_rays
    from dmnerf_b200.render import render_rays
    dev = "cuda"
    wl = synth.workload("dmsr_study")
    lib = _lib.load()
    lib.dmnerf_debug_kprof.restype = C.c_int
    lib.dmnerf_debug_kprof.argtypes = [C.c_void_p, C.c_int]
    fused = "--fused" in sys.argv
    with torch.no_grad():
        if fused:
            coarse, fine, _, _ = make_models(201, 202, 13, dev)
            n = 307200
            ro, rd = torch.from_numpy(wl["rays_o"][:n]).to(dev), torch.from_numpy(wl["rays_d"][:n]).to(dev)
            z = torch.linspace(float(wl["near"]), float(wl["far"]), 64, device=dev)
            for _ in range(2):
                render_rays(ro, rd, coarse, fine, z, want_raw=False, want_samples=False)
            tiles = (n // 2) * 4
        else:
            net = model_from_weights(synth.make_weights(202, 13), dev).eval()
            n = 307200 // 4
            ro, rd = torch.from_numpy(wl["rays_o"][:n]).to(dev), torch.from_numpy(wl["rays_d"][:n]).to(dev)
            z = (torch.rand(n, 192, device=dev).sort(-1).values * 11 + 4).contiguous()
            for _ in range(2):
                mlp_forward_rays(net, ro, rd, z, _lib.IMPL_UMMA)
            tiles = n * 192 // 128
        get_context(dev).sync_check()
    buf = (C.c_longlong * (16 * 148))()
    rc = lib.dmnerf_debug_kprof(buf, 148)
    assert rc == 0, rc
    a = np.frombuffer(buf, dtype=np.int64).reshape(148, 16).astype(np.float64)
    tpc = tiles / 148.0
    print("%s: %d tiles, %.1f per CTA; cycles per tile (mean over CTAs | CTA 0 | CTA 147)" % ("fused" if fused else "mlp", tiles, tpc))
    for i, nm in enumerate(NAMES):
        if nm == "-":
            continue
        print("  %-34s %9.0f | %9.0f | %9.0f" % (nm, a[:, i].mean() / tpc, a[0, i] / tpc, a[147, i] / tpc))
    trace(lib)


def trace(lib):
    import numpy as np
    buf = (C.c_longlong * 256)()
    lib.dmnerf_debug_ktrace.restype = C.c_int
    lib.dmnerf_debug_ktrace.argtypes = [C.c_void_p]
    assert lib.dmnerf_debug_ktrace(buf) == 0
    a = np.frombuffer(buf, dtype=np.int64).reshape(4, 64)
    t0 = a[2][0]
    print("CTA 0, two consecutive tiles: cycles relative to the first acc_full observation")
    print(" step | mma issued (W-stage waits, epilogue waits during the step) | acc_full seen (delta) | epilogue arrived (body)")
    prev = t0
    for i in range(40):
        print("  %3d | %9d (%5d, %5d) | %9d (%6d) | %9d (%5d)" % (i, a[1][i] - t0, a[0][i] >> 32, a[0][i] & 0xffffffff, a[2][i] - t0, a[2][i] - prev,
                                                           a[3][i] - t0, a[3][i] - a[2][i]))
        prev = a[2][i]


if __name__ == "__main__":
    main()
