"""Summarises an .ncu-rep (`ncu --set full`) into a markdown table: duration, DRAM traffic, achieved GB/s, tensor-pipe and
L2 utilisation per launch.   python tools/ncu_summary.py report.ncu-rep > profiles/xxx.md
With `--traffic ID BATCH ALGORITHMIC_BYTES`: also writes profiles/roofline_traffic.json (dram read + write bytes of launch ID),
which bench.py reports as roofline.traffic:
    python tools/ncu_summary.py report.ncu-rep --traffic 0 256 287e6 > profiles/xxx.md"""
import csv
import io
import re
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
col = {c: i for i, c in enumerate(h)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs")]
print("| ID | kernel | " + " | ".join(f"{n} ({units[col[m]]})" if units[col[m]] else n for m, n in want if m in col) + " | HBM GB/s |")
print("|---|---|" + "---|" * (len([1 for m, _ in want if m in col]) + 1))


def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_sec(v, unit):
    return float(v.replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(unit, 1)


traffic = None
if "--traffic" in sys.argv:
    i = sys.argv.index("--traffic")
    traffic = (sys.argv[i + 1], int(sys.argv[i + 2]), float(sys.argv[i + 3]))

for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("b2::", "")[:60]
    vals = [r[col[m]] for m, _ in want if m in col]
    try:
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        t = to_sec(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
        gbs = f"{(rd + wr) / t / 1e9:.0f}"
        if traffic is not None and r[col["ID"]] == traffic[0]:
            import json
            import os
            out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "roofline_traffic.json")
            json.dump({"kernel": name, "batch": traffic[1], "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                       "algorithmic_bytes": traffic[2], "duration_s": t,
                       "source": f"ncu --set full, launch ID {traffic[0]} of {os.path.basename(sys.argv[1])}"}, open(out, "w"), indent=1)
    except Exception:
        gbs = "-"
    print(f"| {r[col['ID']]} | {name} | " + " | ".join(v[:10] for v in vals) + f" | {gbs} |")
