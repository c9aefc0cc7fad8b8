"""Turn gpurun_out/{launches.csv, prof_block.ncu-rep} into tracked summaries under profiles/.
Usage: python tools/summarize_profiles.py r01"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)


def launches():
    p = os.path.join(ROOT, "gpurun_out", "launches.csv")
    if not os.path.exists(p):
        return
    lines = [l for l in open(p) if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        tot[name][0] += 1
        tot[name][1] += v
    S = sum(v[1] for v in tot.values())
    ours = sum(v[1] for k, v in tot.items() if "wb::" in k)
    with open(os.path.join(out, f"{tag}_launch_list_summary.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list of `python bench.py --steps 1 --warmup 3 --rows 32` (window: --launch-skip 3660 -c 1300 = the timed step after the 3 warm-up steps)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised: read SHARES.\n\n")
        f.write(f"total kernel time {S / 1e3:.1f} ms over {sum(v[0] for v in tot.values())} launches; wesep_b200 kernels = {ours / S * 100:.1f} % of it\n\n")
        f.write("| share | avg us | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"| {v[1] / S * 100:.2f} % | {v[1] / v[0]:.1f} | {v[0]} | `{k[:110]}` |\n")
    print("wrote launch list summary")


def full():
    rep = os.path.join(ROOT, "gpurun_out", "prof_block.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "launch__grid_size",
            "launch__block_size", "launch__shared_mem_per_block_dynamic"]
    with open(os.path.join(out, f"{tag}_ncu_full_tcn_block_n32.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` of one Spex+ TCN block fwd+bwd, n=32 rows, "
                "B=256 H=512 K=6399 (tools/profile_block.py 32 8 1)\n\n")
        f.write("Metrics per kernel launch (raw page of the .ncu-rep; the report itself stays in gpurun_out/).\n\n")
        for r in rows[2:]:
            f.write(f"## `{r[idx['Kernel Name']][:100]}`\n\n")
            for w in want:
                if w in idx:
                    f.write(f"- {w}: {r[idx[w]]} {units[idx[w]]}\n")
            f.write("\n")
    print("wrote ncu full summary")


launches()
full()
