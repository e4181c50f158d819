"""Condense gpurun_out/ ncu artefacts into small tracked text files under profiles/ (run here, no GPU needed).
    python tools/summarize_profiles.py r01"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
go = os.path.join(ROOT, "gpurun_out")

# 1. launch lists (gpu__time_duration per launch) -> per-kernel share
LISTS = (("launches.csv", "launches", "python bench.py --steps 1 --warmup 3 --no-cpu-baseline"),
         ("launches_render.csv", "launches_render", "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train (first 80 launches)"),
         ("launches_train.csv", "launches_train", "python tools/train_step.py 3 (three BASELINE-configs[3] training steps: forward + backward + Adam)"))
for lname, oname, cmd in LISTS:
  lp = os.path.join(go, lname)
  if not os.path.exists(lp):
    continue
  if True:
    rows = [r for r in csv.reader(open(lp)) if len(r) > 10]
    hdr = rows[0]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[ik].split("(")[0], [0, 0.0, r[ig], r[ib]])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(out, "%s_%s.md" % (tag, oname)), "w") as f:
        f.write("# ncu launch list of `%s` (gpu__time_duration.sum, "
                "--clock-control none; cold-cache, serialised: compare SHARES)\n\n| kernel | launches | total ms | share | grid | block |\n|---|---|---|---|---|---|\n" % cmd)
        for k, (n, t, g, b) in agg.items():
            f.write("| `%s` | %d | %.3f | %.1f%% | %s | %s |\n" % (k, n, t / 1e6, 100 * t / tot, g, b))
    print("wrote launches summary")

# 2. full-set captures -> key metrics
for rep, fname, title in (("prof_fused.ncu-rep", "ncu_fused_render", "one launch of `mlp_umma_kernel<true>`: the fused render kernel over one "
                           "640x480 frame (307 200 rays, 64+128 samples) as launched by bench.py"),
                          ("prof_umma.ncu-rep", "ncu_fine_mlp", "one launch of `mlp_umma_kernel<false>` (unfused fine network, 37 888 rays x 192 samples)"),
                          ("r02_fused.ncu-rep", "ncu_fused_render", "one launch of `mlp_umma_kernel<true>`: the fused render kernel over one 640x480 "
                           "frame (307 200 rays, 64+128 samples; tools/prof_fused.py 307200)"),
                          ("r02_chain.ncu-rep", "ncu_bwd_chain", "one launch of `bwd_chain_kernel`: the fused gradient chain of the fine network inside a "
                           "training step (196 608 samples; tools/train_step.py)"),
                          ("r02_fwdtrain.ncu-rep", "ncu_train_forward", "one launch of `mlp_umma_kernel<false>` with activations kept: the fine network's "
                           "training forward (196 608 samples; tools/train_step.py)"),
                          ("r02_dw.ncu-rep", "ncu_dw_gemm", "one BATCHED launch of `gemm_tn_tc_kernel<256,256>`: the eight [256x256] weight-gradient "
                           "products of the network whose backward runs first (144 CTAs = 8 products x 18 sample slices) inside a training step"),
                          ("r02_dw_b.ncu-rep", "ncu_dw_gemm_b", "the same batched launch for the other network of the step")):
  rp = os.path.join(go, rep)
  if os.path.exists(rp):
      raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
      rows = list(csv.reader(raw.splitlines()))
      hdr, units, vals = rows[0], rows[1], rows[2]
      want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
              "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex.sum",
              "lts__t_sectors_srcunit_tex.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
              "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
              "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__inst_executed.sum"]
      got = {}
      for h, u, v in zip(hdr, units, vals):
          for w in want:
              if h == w or h.endswith("." + w):
                  got[w] = (v, u)
      with open(os.path.join(out, "%s_%s.md" % (tag, fname)), "w") as f:
          f.write("# ncu --set full, " + title + "\n\n"
                  "Captured with `tools/gpu_prof.sh` (`--clock-control none`); the .ncu-rep stays in gpurun_out/.\n\n"
                  "| metric | value | unit |\n|---|---|---|\n")
          for w in want:
              if w in got:
                  f.write("| %s | %s | %s |\n" % (w, got[w][0], got[w][1]))
          try:
              scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
              dr = float(got["dram__bytes_read.sum"][0].replace(",", "")) * scale[got["dram__bytes_read.sum"][1]]
              dw = float(got["dram__bytes_write.sum"][0].replace(",", "")) * scale[got["dram__bytes_write.sum"][1]]
              if fname == "ncu_fused_render":
                  json.dump({"kernel": "mlp_umma_kernel<true> (fused render, one 640x480 frame)", "dram_bytes_per_launch": dr + dw,
                             "dram_read_bytes": dr, "dram_write_bytes": dw, "source": "%s_%s.md" % (tag, fname)},
                            open(os.path.join(out, "traffic.json"), "w"), indent=1)
              f.write("\nDRAM traffic per launch: read %s %s + write %s %s.\n" % (got["dram__bytes_read.sum"] + got["dram__bytes_write.sum"]))
          except Exception:
              pass
      print("wrote ncu summary")
