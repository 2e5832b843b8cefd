"""Per-CUDA-source-line summary of an ncu report (samples, instructions executed, top stall reasons).
Usage: python tools/ncu_lines.py <report.ncu-rep> [kernel-id] [top-n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 2 and sys.argv[2] != "-":
    cmd += ["--kernel-id", sys.argv[2]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
fname = ""
data = []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        s, n = int(r[si]), int(r[ie])
    except ValueError:
        continue
    st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:3]
    data.append((s, n, fname, r[0], r[1], st))
ts, ti = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
print(f"total samples {ts}  warp instructions {ti}")
for s, n, f, ln, src, st in sorted(data, key=lambda d: -d[0])[:top]:
    stall = " ".join(f"{h}:{c}" for c, h in st if c)
    print(f"{s:6d} {100*s/ts:5.1f}% | {n:9d} {100*n/ti:5.1f}% | {f}:{ln:>4} {src.strip()[:90]}  [{stall}]")
