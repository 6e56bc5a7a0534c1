"""DRAM traffic per kernel from an ncu metrics CSV of the bench command -> profiles/rNN_dram_traffic_<config>.json (bench.py's
`roofline.traffic`).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16|attn_ \
        -s 200 -c 100 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
    python tools/ncu_traffic.py gpurun_out/traffic.csv "<source command>" > profiles/r02_dram_traffic_vit_b16.json
"""
import collections
import csv
import json
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"^void |vb::<unnamed>::|<unnamed>::|\(.*", "", r[ix["Kernel Name"]]).replace("unnamed>::", "")
    k = per.setdefault((r[ix["ID"]], name), {})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    m = r[ix["Metric Name"]]
    if m.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    elif m.startswith("gpu__time"):
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(unit, 1)
    k[m] = v
agg = collections.OrderedDict()
for (_, name), k in per.items():
    a = agg.setdefault(name, dict(launches=0, rd=0.0, wr=0.0, us=0.0))
    a["launches"] += 1
    a["rd"] += k.get("dram__bytes_read.sum", 0.0)
    a["wr"] += k.get("dram__bytes_write.sum", 0.0)
    a["us"] += k.get("gpu__time_duration.sum", 0.0)
out = dict(source=sys.argv[2] if len(sys.argv) > 2 else "", per_kernel={})
g = dict(launches=0, bytes=0.0)
for name, a in agg.items():
    out["per_kernel"][name] = dict(launches=a["launches"], dram_read_MB_per_launch=round(a["rd"] / a["launches"] / 1e6, 1),
                                   dram_write_MB_per_launch=round(a["wr"] / a["launches"] / 1e6, 1),
                                   avg_us_under_ncu=round(a["us"] / a["launches"], 1))
    if name.startswith("gemm_bf16_kernel"):
        g["launches"] += a["launches"]
        g["bytes"] += a["rd"] + a["wr"]
out["gemm_class"] = dict(launches=g["launches"], dram_bytes_per_launch=(g["bytes"] / g["launches"]) if g["launches"] else None)
print(json.dumps(out, indent=1))
