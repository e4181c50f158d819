"""SASS evidence for the tensor-core / TMEM / TMA-engine claims: per-kernel counts of the Blackwell mnemonics in the shipped
library (cuobjdump -sass).        python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dm-nerf_b200", "lib", "libdmnerf_b200.so")
WATCH = ("UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "UBLKRED", "UTMALDG", "UTMASTG", "UTMAREDG", "SYNCS", "HMMA", "FFMA", "MUFU",
         "REDG", "ATOMG", "RED", "LDG", "STG", "LDS", "STS", "SHFL", "BAR")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    cur[w] += 1
    demangled = {}
    try:
        names = "\n".join(kernels)
        dm = subprocess.run(["cu++filt"], input=names, capture_output=True, text=True).stdout.splitlines()
        demangled = dict(zip(kernels, dm))
    except Exception:
        pass
    print("# cuobjdump -sass %s : instruction counts per kernel (sm_100a)" % os.path.relpath(LIB, ROOT))
    print("# tcgen05.mma = UTCHMMA, tcgen05.commit = UTCBAR, tcgen05.ld/st = LDTM/STTM, cp.async.bulk (TMA engine, 1-D) = UBLKCP,")
    print("# tensor-map TMA would be UTMALDG/UTMASTG (not used: the weight image is pre-swizzled, a 1-D bulk copy is exact)")
    cols = [w for w in WATCH if any(c[w] for c in kernels.values())]
    print("%-86s %7s " % ("kernel", "instrs") + " ".join("%7s" % c[:7] for c in cols))
    for k, c in kernels.items():
        name = re.sub(r"\(.*", "", demangled.get(k, k))[:86]
        print("%-86s %7d " % (name, c["_total"]) + " ".join("%7d" % c[w] for w in cols))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print("%-86s %7d " % ("TOTAL", tot["_total"]) + " ".join("%7d" % tot[w] for w in cols))


if __name__ == "__main__":
    sys.exit(main())
