"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/<tag>_launches.csv  > profiles/<tag>_launches.md
    python tools/ncu_summary.py full gpurun_out/<tag>_full_<kernel>.ncu-rep  > profiles/<tag>_full_<kernel>.md

`launches`: per-kernel launch count, total / average gpu__time_duration and share of the captured window
(the per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes).
`full`: the handful of raw metrics the roofline discussion uses (duration, DRAM bytes, DRAM / L2 / tensor-pipe
utilisation, registers, occupancy) for every captured launch.
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "derived__lts__lts2xbar_bytes.sum.per_second",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
]


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", row["Kernel Name"]))
        name = name.replace("d3d::<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
        a = agg.setdefault(name, [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list: {path}\n")
    print(f"{n} launches captured, {tot:.1f} ms of kernel time (serialised, cold-cache: compare shares)\n")
    print("| kernel | launches | total ms | avg ms | share | grid | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:70]}` | {v[0]} | {v[1]:.3f} | {v[1] / v[0]:.4f} | {100 * v[1] / tot:.1f}% | {v[2]} | {v[3]} |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full: {path}\n")
    names = [re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("d3d::<unnamed>::", "")[:60] for r in data]
    print("| metric | unit | " + " | ".join(f"`{n}`" for n in names) + " |")
    print("|---|---|" + "---:|" * len(data))
    for k in KEYS:
        if k in idx:
            print(f"| {k} | {units[idx[k]]} | " + " | ".join(r[idx[k]] for r in data) + " |")


def traffic(path, gemm, tokens):
    """profiles/gemm_traffic.json: average DRAM bytes per GEMM launch from a capture of the four GEMMs of one block."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def to_bytes(v, u):
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]

    per = []
    for r in data:
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])[-40:]
        per.append({"kernel": name, "dram_read": rd, "dram_write": wr,
                    "ms": float(r[idx["gpu__time_duration.sum"]].replace(",", ""))})
    avg = sum(p["dram_read"] + p["dram_write"] for p in per) / len(per)
    # algorithmic HBM bytes per token: qkv 6144, proj 6144, fc1 6144, fc2 8192
    algo = int(tokens) * (6144 + 6144 + 6144 + 8192) / 4
    print(json.dumps({"gemm": gemm, "tokens_per_launch": int(tokens), "dram_bytes_per_launch_avg": avg,
                      "algorithmic_bytes_per_launch_avg": algo, "launches": per,
                      "source": f"ncu --set full, {path.split('/')[-1]}"}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
