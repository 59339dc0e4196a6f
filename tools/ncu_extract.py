"""Turn the `ncu --set full` captures of a GPU session into profiles/ncu_traffic.json (what bench.py's roofline.traffic reads)
and a markdown summary.   python tools/ncu_extract.py TAG [steps_in_persistent_capture=300]"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1]
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 300
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        try:
            d[h] = float(v.replace(",", "")) * UNIT.get(u, 1.0)
        except ValueError:
            d[h] = v
    return d


def sha16(name):
    return hashlib.sha256(open(os.path.join(ROOT, "distributed_sgd_b200", "csrc", name), "rb").read()).hexdigest()[:16]


traffic, lines = {}, [f"# ncu --set full captures of session `{TAG}` (one launch each, --clock-control none)\n"]
for key, rep, src, note in (("k_sync_persistent", f"{TAG}_prof_persist", "dsgd_persistent.cuh", f"{STEPS} SGD steps of batch 256 in one launch"),
                            ("k_stream_rows_eval", f"{TAG}_prof_stream_eval", "dsgd_stream.cuh", "evaluation pass over the 560 000 train rows"),
                            ("k_stream_rows_scatter", f"{TAG}_prof_stream_scatter", "dsgd_stream.cuh",
                             "gradient of 262 144 random rows on trained weights")):
    path = os.path.join(ROOT, "gpurun_out", rep + ".ncu-rep")
    if not os.path.exists(path):
        lines.append(f"\n## {key}: no capture ({rep}.ncu-rep missing)\n")
        continue
    d = raw(path)
    dram = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
    rec = {"source": src, "source_sha16": sha16(src), "capture": f"gpurun_out/{rep}.ncu-rep ({note})", "dram_bytes": dram,
           "duration_s": d["gpu__time_duration.sum"]}
    if key == "k_sync_persistent":
        rec["dram_bytes_per_sgd_step_batch256"] = dram / STEPS
    traffic[key] = rec
    lines.append(f"\n## {d.get('Kernel Name', key)}\n{note}\n\n| metric | value |\n|---|---|")
    for m in WANT:
        if m in d:
            v = d[m]
            lines.append(f"| {m} | {v:.6g} |" if isinstance(v, float) else f"| {m} | {v} |")
    lines.append(f"| dram bytes read + written | {dram / 1e6:.2f} MB |")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)
with open(os.path.join(ROOT, "profiles", f"{TAG}_ncu_summary.md"), "w") as f:
    f.write("\n".join(lines) + "\n")
print("\n".join(lines))
