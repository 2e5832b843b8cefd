"""Warp-state samples of the fused kernel split by WARP ROLE, from an ncu report with SASS-level sampling data.

The control warps (TMA producers, MMA issuers, the flag agent) and the compute warpgroups run disjoint SASS regions of
tc_layer_kernel: the control branch starts at `USETMAXREG.DEALLOC` (setmaxnreg.dec) and the compute part at
`USETMAXREG.TRY_ALLOC` (setmaxnreg.inc); everything before the first of the two is the common prologue.  Per role: samples,
share, the top stall reasons, and the same split for the mbarrier try_wait instructions (SYNCS.PHASECHK...), whose source
line (umma.cuh) is shared by both roles.

Usage: python tools/ncu_roles.py <report.ncu-rep> [kernel-id]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]
if len(sys.argv) > 2:
    cmd += ["--kernel-id", sys.argv[2]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = [r for r in csv.reader(out.splitlines()) if len(r) > 8]
hdr = next(r for r in rows if "# Samples" in r and "Source" in r)
isrc, isam = hdr.index("Source"), hdr.index("# Samples")
stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows:
    if r is hdr or len(r) < len(hdr):
        continue
    try:
        s = int(r[isam])
    except ValueError:
        continue
    data.append((r[isrc], s, [int(r[i] or 0) for i, _ in stall]))
dec = next((i for i, d in enumerate(data) if "USETMAXREG" in d[0] and "DEALLOC" in d[0]), None)          # setmaxnreg.dec
inc = next((i for i, d in enumerate(data) if "USETMAXREG" in d[0] and "TRY_ALLOC" in d[0]), None)        # setmaxnreg.inc
if dec is None or inc is None:
    sys.exit("setmaxnreg markers not found in the SASS listing")
first = min(dec, inc)


def role(i):
    if i < first:
        return "common prologue"
    if dec < inc:
        return "control warps (TMA / MMA issue / agent)" if i < inc else "compute warpgroups"
    return "compute warpgroups" if i < dec else "control warps (TMA / MMA issue / agent)"


tot = sum(d[1] for d in data) or 1
agg = {}
top_instr = {}
for i, (src, s, st) in enumerate(data):
    top_instr.setdefault(role(i), []).append((s, " ".join(src.split())[:70], max(((v, stall[k][1]) for k, v in enumerate(st)), default=(0, ""))[1]))
    a = agg.setdefault(role(i), [0, [0] * len(stall), 0])
    a[0] += s
    for k, v in enumerate(st):
        a[1][k] += v
    if "SYNCS" in src and "TRYWAIT" in src.upper().replace("_", "").replace(".", "") or "SYNCS.PHASECHK" in src or "NANOSLEEP.SYNCS" in src:
        a[2] += s
print(f"total warp-state samples {tot}  (SASS instructions {len(data)}, setmaxnreg.dec @{dec}, setmaxnreg.inc @{inc})")
for name, (s, st, sw) in agg.items():
    top = sorted(((v, stall[k][1]) for k, v in enumerate(st)), reverse=True)[:6]
    print(f"{name:42s} {s:8d} samples {100 * s / tot:5.1f} %   of which in mbarrier waits (SYNCS.PHASECHK / NANOSLEEP.SYNCS) {sw} ({100 * sw / max(s, 1):.1f} %)")
    print("    " + "  ".join(f"{n}:{v} ({100 * v / max(s, 1):.0f}%)" for v, n in top if v))
    for smp, txt, why in sorted(top_instr[name], reverse=True)[:8]:
        print(f"      {smp:7d} {100 * smp / max(s, 1):5.1f} %  {txt:70s} [{why}]")
