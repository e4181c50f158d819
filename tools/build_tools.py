"""Build the stand-alone hardware probes under tools/ into tools/bin/ (git-ignored, shipped by gpurun)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "bin")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"]


def main():
    os.makedirs(BIN, exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    for src in sorted(f for f in os.listdir(HERE) if f.endswith(".cu")):
        out = os.path.join(BIN, src[:-3])
        if os.path.exists(out) and os.path.getmtime(out) >= max(
                os.path.getmtime(os.path.join(HERE, src)),
                os.path.getmtime(os.path.join(HERE, "..", "dm-nerf_b200", "csrc", "umma.cuh"))):
            continue
        r = subprocess.run([nvcc] + FLAGS + [os.path.join(HERE, src), "-o", out], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout)
            raise RuntimeError("nvcc failed on tools/%s" % src)


if __name__ == "__main__":
    main()
