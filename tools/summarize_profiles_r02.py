"""Turn the round-2 evidence in gpurun_out/ (tools/gpu_round2.sh) into tracked summaries under profiles/.
Usage: python tools/summarize_profiles_r02.py"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
G = os.path.join(ROOT, "gpurun_out")
OUT = os.path.join(ROOT, "profiles")
tag = "r02"


def launches():
    p = os.path.join(G, "r02_launches_spex.csv")
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
    with open(os.path.join(OUT, f"{tag}_launch_list_spex_summary.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pbsrnn` "
                "(window: --launch-skip 3000 -c 1400 = the end of the warm-up and the timed Spex+ step)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised: read SHARES "
                "(warm, overlapped-as-run times: r02_kernel_times_spex_n32.md).\n\n")
        f.write(f"total kernel time {S / 1e3:.1f} ms over {sum(v[0] for v in tot.values())} launches; wesep_b200 kernels = {ours / S * 100:.1f} % of it\n\n")
        f.write("| share | avg us | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:32]:
            f.write(f"| {v[1] / S * 100:.2f} % | {v[1] / v[0]:.1f} | {v[0]} | `{k[:110]}` |\n")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def full(rep, outname, title):
    rep = os.path.join(G, rep)
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(OUT, outname), "w") as f:
        f.write(f"# {tag}: {title}\n\nMetrics per kernel launch (raw page of the .ncu-rep; the report itself stays in gpurun_out/).\n\n")
        for r in rows[2:]:
            f.write(f"## `{r[idx['Kernel Name']][:100]}`\n\n")
            for w in WANT:
                if w in idx:
                    f.write(f"- {w}: {r[idx[w]]} {units[idx[w]]}\n")
            f.write("\n")


def sass():
    o = os.path.join(ROOT, "wesep_b200", "csrc", "build", "lstm_rec.o")
    if not os.path.exists(o):
        return
    txt = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
    cur, cnt = None, collections.defaultdict(lambda: collections.Counter())
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            continue
        for key in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "UTCATOMSWS", "MUFU.EX2", "MUFU.RCP", "ST.E.128", "STS.128", "BAR.ARV", "BAR.SYNC", "ELECT"):
            if cur and re.search(r"\b" + re.escape(key) + r"\b", ln):
                cnt[cur][key] += 1
    with open(os.path.join(OUT, f"{tag}_sass_lstm_rec.md"), "w") as f:
        f.write(f"# {tag}: SASS evidence for the persistent BLSTM recurrence kernels (`cuobjdump -sass wesep_b200/csrc/build/lstm_rec.o`)\n\n"
                "tcgen05.mma -> UTCHMMA (A operand from tensor memory: `tmem[..]`), tcgen05.commit -> UTCBAR, tcgen05.ld / st -> LDTM / STTM, "
                "cp.async.bulk shared::cta -> shared::cluster -> UBLKCP.S.S, mbarrier -> SYNCS.  One ELECT (elect.sync) guards the whole MMA loop: "
                "the UTCHMMA instructions are emitted back to back.\n\n| kernel | " + " | ".join(
                    ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "MUFU.EX2", "MUFU.RCP", "ELECT"]) + " |\n|---|" + "---:|" * 9 + "\n")
        for k in sorted(cnt):
            if "8E" in k or "Li8" in k:
                f.write(f"| `{k[:60]}` | " + " | ".join(str(cnt[k][x]) for x in ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "MUFU.EX2", "MUFU.RCP", "ELECT"]) + " |\n")


os.makedirs(OUT, exist_ok=True)
launches()
full("r02_prof_lstm_rec.ncu-rep", "r02_ncu_full_lstm_rec.md",
     "`ncu --set full --clock-control none --import-source on -k regex:lstm_rec` of the persistent BLSTM recurrence at the band_rnn shape of "
     "BASELINE config 3 (S = 501 steps, Q = 512 sequences, Hd = 256; `tools/run_lstm_rec_once.py 501 512 256 128`)")
full("r02_prof_block.ncu-rep", "r02_ncu_full_tcn_block_n32.md",
     "`ncu --set full --clock-control none --import-source on` of one Spex+ TCN block fwd+bwd, n=32 rows, B=256 H=512 K=6399 "
     "(tools/profile_block.py 32 8 1); 2-CTA GEMMs in the mixed tf32 + bf16 mode (PRO 2 kernel in 3xTF32)")
sass()
for src, dst in (("r02_bench.log", "r02_bench_1gpu.json"), ("r02_bench_ref.log", "r02_bench_reference_arm.json"),
                 ("r02_kernel_times_spex_n32.md", "r02_kernel_times_spex_n32.md"),
                 ("r02_kernel_times_pbsrnn_n16.md", "r02_kernel_times_pbsrnn_n16.md")):
    p = os.path.join(G, src)
    if os.path.exists(p):
        if dst.endswith(".json"):
            line = open(p).read().strip().splitlines()[-1]
            json.loads(line)
            open(os.path.join(OUT, dst), "w").write(line + "\n")
        else:
            shutil.copy(p, os.path.join(OUT, dst))
print(sorted(os.listdir(OUT)))
