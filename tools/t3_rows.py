"""Summarise an ncu metric pass over the kernels of one decoder step of the tf32 x 3 path (precision="fp32") at cfg3 size:
one row per launch of a layer (+ the step's first / last GEMM), with duration, tensor-pipe share, DRAM bytes and achieved GB/s, and the
algorithmic TFLOP/s (3 tf32 MMAs per product counted as ONE product, i.e. fp32-equivalent work).
Usage: python tools/t3_rows.py <ncu.csv> [rows=204800]   (csv from: ncu --metrics ... -k regex:t3_ --csv, tools/gpu_final.sh)"""
import csv
import sys
from collections import OrderedDict

rows_n = int(sys.argv[2]) if len(sys.argv) > 2 else 204800
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ik, im, iv, ig, igrid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID"), h.index("Grid Size")
d = OrderedDict()
for r in rows[1:]:
    name = r[ik].split("(")[0].replace("void ", "").strip()
    d.setdefault((int(r[ig]), name, r[igrid]), {})[r[im]] = float(r[iv].replace(",", ""))
H, FFN, M = 160, 320, 80
# name by (kernel, grid): the launches of one layer in order
names = {("t3_gemm_kernel", "(1600, 3, 1)"): ("qkv GEMM  (AdaRMSNorm prologue, K 160 -> N 480)", 2 * H * 3 * H),
         ("t3_gemm_kernel", "(1600, 4, 1)"): ("ffn0 GEMM (AdaRMSNorm prologue, SwiGLU epilogue, 160 -> 640 -> 320)", 2 * H * 2 * FFN),
         ("t3_attn_kernel", "(7, 4, 256)"): None,
         ("t3_rowstats_kernel", None): ("row statistics of a norm prologue", 0),
         ("t3_kvimg_kernel", None): ("context K | V operand images of one layer (once per generate)", 0),
         ("t3_pack_jobs_kernel", None): ("weight images of the step (62 blocks)", 0)}
print(f"{'kernel':74s} {'us':>8s} {'tensor %':>8s} {'issue %':>8s} {'DRAM MB':>9s} {'GB/s':>7s} {'% HBM':>6s} {'TFLOP/s':>8s}")
seen = {}
for (kid, kname, grid), v in d.items():
    t_us = v["gpu__time_duration.sum"] / 1e3
    mb = (v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]) / 1e6
    label, flop_row = None, 0
    if kname.startswith("t3_attn_kernel"):
        # <0>: the softmax warps stage k | v themselves (band attention); <1>: ready-made context K | V operand images (cross-attention)
        win = "<1>" not in kname
        label = "window attention (band +-64, 4 heads)" if win else "cross attention (400 context tokens, 4 heads)"
        flop_row = (4 * 129 * 40 * 2 * 2) if win else (4 * 400 * 40 * 2 * 2)
    elif kname == "t3_gemm_kernel" and grid == "(1600, 1, 1)":
        n = seen.get("g1", 0)
        seen["g1"] = n + 1
        label = "GEMM N <= 160, one block column (proj / q_proj / out / ffn3 / in_proj / out_proj, in launch order)"
        flop_row = 2 * H * H
    else:
        for (kn, gr), val in names.items():
            if kn == kname and (gr is None or gr == grid) and val:
                label, flop_row = val
    if label is None:
        label = f"{kname} {grid}"
    tf = flop_row * rows_n / (t_us * 1e-6) / 1e12 if flop_row else 0.0
    print(f"{label[:74]:74s} {t_us:8.1f} {v.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):8.1f} "
          f"{v.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):8.1f} {mb:9.1f} {mb / t_us * 1e3:7.0f} "
          f"{100 * mb / t_us * 1e3 / 6547.8:6.1f} {tf:8.1f}")
