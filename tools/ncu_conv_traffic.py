"""Turn an ncu CSV of one training step (metrics dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum;
see profiles/README) into the per-launch DRAM traffic of the convolution kernels that bench.py reports as
roofline.traffic.   python tools/ncu_conv_traffic.py gpurun_out/ncu_step_dram.csv profiles/r1_conv_dram.json"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.defaultdict(lambda: collections.defaultdict(float))   # launch id -> metric -> value
names = {}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
         "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        continue
    v = float(r[ix["Metric Value"]].replace(",", "")) * scale.get(r[ix["Metric Unit"]], 1.0)
    per[r[ix["ID"]]][r[ix["Metric Name"]]] = v
    names[r[ix["ID"]]] = r[ix["Kernel Name"]].split("(")[0]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i, m in per.items():
    a = agg[names[i]]
    a[0] += 1
    a[1] += m.get("dram__bytes_read.sum", 0.0)
    a[2] += m.get("dram__bytes_write.sum", 0.0)
    a[3] += m.get("gpu__time_duration.sum", 0.0)
conv = [v for k, v in agg.items() if "conv_fwd_kernel" in k or "conv_wgrad_kernel" in k]
n = sum(v[0] for v in conv)
out = {
    "source": sys.argv[1],
    "conv_launches": n,
    "dram_bytes_per_launch": sum(v[1] + v[2] for v in conv) / max(n, 1),
    "dram_bytes_per_step_conv": sum(v[1] + v[2] for v in conv),
    "kernels": {k: {"launches": v[0], "dram_read_bytes": v[1], "dram_write_bytes": v[2], "time_s": v[3]}
                for k, v in sorted(agg.items(), key=lambda kv: -kv[1][3])[:25]},
}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: out[k] for k in ("conv_launches", "dram_bytes_per_launch", "dram_bytes_per_step_conv")}))
