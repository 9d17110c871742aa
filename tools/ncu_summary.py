"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/ncu_summary.py report <file.ncu-rep> <out.csv>     # selected metrics of every captured launch
    python tools/ncu_summary.py traffic <launches.csv> <regex> <out.json>
        # average dram__bytes_read+write per launch of the kernels matching <regex> in a launch list captured with
        # --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum  (bench.py reads the json)
"""
import csv
import io
import json
import re
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def report(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in idx])
        for r in rows[2:]:
            w.writerow([r[i][:120] for i in idx])
    print(f"{out}: {len(rows) - 2} launches")


def traffic(path, pattern, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = {}
    for r in csv.DictReader(lines):
        if not re.search(pattern, r["Kernel Name"]):
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
        d = per.setdefault(r["ID"], {})
        d[r["Metric Name"]] = v * mult
    n = len(per)
    rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in per.values())
    wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in per.values())
    ns = sum(d.get("gpu__time_duration.sum", 0.0) for d in per.values())
    res = {"kernel_regex": pattern, "launches": n, "dram_bytes_per_launch": (rd + wr) / max(1, n),
           "dram_read_bytes_per_launch": rd / max(1, n), "dram_write_bytes_per_launch": wr / max(1, n),
           "avg_launch_us_under_ncu": ns / max(1, n) / 1e3, "source": path}
    json.dump(res, open(out, "w"), indent=1)
    print(res)


if __name__ == "__main__":
    if sys.argv[1] == "report":
        report(sys.argv[2], sys.argv[3])
    else:
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
