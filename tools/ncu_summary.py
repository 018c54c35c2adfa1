"""Summary of an ncu report (CPU only): per launch the duration, the DMMA pipe activity, DRAM bytes and occupancy facts.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [note] > profiles/x.json      (development aid)"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]


def col(name):
    return hdr.index(name) if name in hdr else None


def num(r, name, scale=1.0):
    i = col(name)
    if i is None or r[i] == "":
        return None
    v = float(r[i].replace(",", ""))
    u = units[i]
    if u == "Kbyte":
        v *= 1e3
    elif u == "Mbyte":
        v *= 1e6
    elif u == "Gbyte":
        v *= 1e9
    elif u == "ms":
        v *= 1e3          # durations in us
    elif u == "ns":
        v *= 1e-3
    elif u == "s":
        v *= 1e6
    return v * scale


out = {"report": rep, "note": note, "launches": []}
for r in rows[2:]:
    name = r[col("Kernel Name")].replace("<unnamed>::", "").replace("void ", "")
    rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
    out["launches"].append({
        "kernel": name.split("(")[0],
        "grid": r[col("Grid Size")] if col("Grid Size") is not None else None,
        "block": r[col("Block Size")] if col("Block Size") is not None else None,
        "duration_us": num(r, "gpu__time_duration.sum"),
        "dmma_pipe_pct_active": num(r, "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_pct_active": num(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "dram_read_MB": rd / 1e6 if rd is not None else None,
        "dram_write_MB": wr / 1e6 if wr is not None else None,
        "dram_throughput_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "l2_bytes_MB": (num(r, "lts__t_bytes.sum") or 0) / 1e6,
        "registers": num(r, "launch__registers_per_thread"),
        "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "sm_throughput_pct": num(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "smem_bank_conflicts": num(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    })
print(json.dumps(out, indent=1))
