"""Text summary of an `ncu --set full --import-source on` report for profiles/: the headline metrics of the profiled
kernel (raw page) and the SASS lines that collect the most warp-stall samples (source page).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-name-substring] > profiles/x.txt
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "smsp__inst_executed.sum"]


def run(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(run("raw"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
if not hi:
    sys.exit("no raw page in %s" % rep)
hdr, units = rows[hi[0]], rows[hi[0] + 1]
print("report:", rep)
for r in rows[hi[0] + 2:]:
    if len(r) != len(hdr) or (want and want not in r[hdr.index("Kernel Name")]):
        continue
    print("kernel:", r[hdr.index("Kernel Name")][:100], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for k in KEYS:
        for j, h in enumerate(hdr):
            if h == k or h.endswith("." + k):
                print("  %-70s %s %s" % (k, r[j], units[j]))
                break
src = list(csv.reader(io.StringIO(run("source"))))
hi = [i for i, r in enumerate(src) if r and "Source" in r and any("Sampling" in c or "Samples" in c for c in r)]
if hi:
    h = src[hi[0]]
    cs = h.index("Source")
    cand = [i for i, c in enumerate(h) if c.strip() in ("# Samples", "Warp Stall Sampling (All Samples)", "Warp Stall Sampling (All Cycles)")]
    if cand:
        cn = cand[0]
        body = [r for r in src[hi[0] + 1:] if len(r) == len(h)]

        def num(v):
            try:
                return float(v.replace(",", ""))
            except ValueError:
                return 0.0
        tot = sum(num(r[cn]) for r in body)
        print("warp-stall samples: %d; top SASS lines:" % tot)
        for r in sorted(body, key=lambda r: -num(r[cn]))[:25]:
            print("  %6d %5.1f%%  %s" % (num(r[cn]), 100 * num(r[cn]) / max(tot, 1), r[cs][:90]))
