"""tools/traffic_json.py -- ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum of the tensor-core conv launches of one
step) -> profiles/r03_traffic_conv_tc.json, which bench.py reads for roofline.traffic.
usage: python tools/traffic_json.py gpurun_out/traffic.csv 1024 > profiles/r03_traffic_conv_tc.json"""
import csv
import json
import sys

path, batch = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = {}
for r in rows[1:]:
    if len(r) != len(hdr) or not r[ix["ID"]].isdigit():
        continue
    unit, val = r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0, "nsecond": 1e-6, "second": 1e3}.get(unit, 1.0)
    per.setdefault(int(r[ix["ID"]]), {})[r[ix["Metric Name"]]] = val * scale
n = len(per)
rd = sum(v.get("dram__bytes_read.sum", 0.0) for v in per.values())
wr = sum(v.get("dram__bytes_write.sum", 0.0) for v in per.values())
ms = sum(v.get("gpu__time_duration.sum", 0.0) for v in per.values())
print(json.dumps({
    "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_conv_tc (tools/one_step.py %d 2, MARS_GRAPH=0: the tensor-core conv launches of the second pass)" % batch,
    "batch_per_gpu": batch, "kernel_group": "conv_tcgen05_i8", "launches_per_step": n,
    "dram_bytes_per_step": rd + wr, "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr,
    "dram_bytes_per_launch_avg": (rd + wr) / max(n, 1), "dram_bytes_per_image": (rd + wr) / batch, "time_ms_per_step_under_ncu": ms}, indent=1))
