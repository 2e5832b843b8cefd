#!/bin/bash
# ncu rows (time, DRAM bytes, pipes) of the kernels behind one tools/kernel_rows.py case.  Usage: bash tools/gpu_ncu_only.sh <substr> <tag>
set -u
PAT=$1; TAG=${2:-x}; OUT=gpurun_out; mkdir -p $OUT
python tools/kernel_rows.py --only "$PAT" 2>&1 | grep -v "^\[build\|^JSON"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file $OUT/only_${TAG}.csv python tools/kernel_rows.py --once --only "$PAT" > $OUT/only_${TAG}.log 2>&1
python - <<P
import csv
rows=[r for r in csv.reader(open("$OUT/only_${TAG}.csv")) if len(r)>10]
h=rows[0]; ik,im,iv,ii=h.index("Kernel Name"),h.index("Metric Name"),h.index("Metric Value"),h.index("ID")
by={}
for r in rows[1:]: by.setdefault((int(r[ii]),r[ik][:50]),{})[r[im].split(".")[0].replace("sm__","").replace("smsp__","")]=float(r[iv].replace(",",""))
for k,m in sorted(by.items()):
    if "at::" in k[1] or "pack" in k[1]: continue
    ns=m["gpu__time_duration"]; b=m["dram__bytes_read"]+m["dram__bytes_write"]
    print(f"{k[1]:50s} {ns/1e3:8.1f} us {b/1e6:8.1f} MB {b/ns:7.1f} GB/s fma {m.get('pipe_fma_cycles_active',0):5.1f} xu {m.get('inst_executed_pipe_xu',0):5.1f} tensor {m.get('pipe_tensor_cycles_active',0):5.1f} issue {m.get('issue_active',0):5.1f} warps {m.get('warps_active',0):5.1f}")
P
