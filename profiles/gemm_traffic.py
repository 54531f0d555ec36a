"""DRAM traffic of the tcgen05 GEMM launches of one train step, from an ncu CSV:

    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none -k regex:gemm_tcgen05 --csv --log-file gpurun_out/gemm_traffic.csv \
        python profiles/profile_step.py --config xl2
    python profiles/gemm_traffic.py gpurun_out/gemm_traffic.csv profiles/r01_gemm_traffic.json

bench.py copies `dram_bytes_per_launch` into roofline.traffic (per launch, like roofline.achieved).
"""
import collections
import csv
import json
import sys

src, dst = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if not l.startswith("==")]
per = collections.defaultdict(dict)
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3,
             "usecond": 1, "msecond": 1e3}.get(u, 1)
    per[r["ID"]][r["Metric Name"]] = v * scale
n = len(per)
rd = sum(p.get("dram__bytes_read.sum", 0) for p in per.values())
wr = sum(p.get("dram__bytes_write.sum", 0) for p in per.values())
us = sum(p.get("gpu__time_duration.sum", 0) for p in per.values())
out = {"kernel": "gemm_tcgen05_kernel", "launches": n, "dram_read_bytes": rd, "dram_write_bytes": wr,
       "dram_bytes_per_launch": (rd + wr) / max(n, 1), "time_us": us,
       "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over every tcgen05 GEMM launch of one SiT-XL/2 train step "
                 "(profiles/profile_step.py, local batch 32)"}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out))
