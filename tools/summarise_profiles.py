"""Turn the ncu reports in gpurun_out/ into the tracked summaries under profiles/ (run in the build container).

    python tools/summarise_profiles.py r01
"""
import csv
import io
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(REPO, "gpurun_out"), os.path.join(REPO, "profiles")
os.makedirs(P, exist_ok=True)
KEEP = ("Duration", "Elapsed Cycles", "SM Frequency", "Executed Ipc Active", "Issue Slots Busy", "SM Busy",
        "Registers Per Thread", "Block Size", "Grid Size", "Cluster Size", "Dynamic Shared Memory Per Block", "Static Shared Memory Per Block",
        "Theoretical Occupancy", "Achieved Occupancy", "Block Limit", "Waves Per SM", "DRAM Throughput", "Memory Throughput",
        "L1/TEX Hit Rate", "L2 Hit Rate", "Mem Busy", "Eligible Warps", "Issued Warp", "No Eligible", "Avg. Active Threads",
        "Compute (SM) Throughput", "Local", "Shared Memory Configuration")
METRICS = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "smsp__inst_executed.sum",
           "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
           "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.sum",
           "smsp__inst_executed_pipe_fma.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__cluster_dim_x", "launch__grid_size", "launch__block_size")
traffic = {}
for cfg in (2, 3, 4, 5):
    rep = os.path.join(G, f"full_c{cfg}.ncu-rep")
    if not os.path.exists(rep):
        continue
    cmdf = os.path.join(G, f"full_c{cfg}.cmd")
    cmd = open(cmdf).read().strip() if os.path.exists(cmdf) else f"python tools/prof_run.py --config {cfg} --iters 40{' --objects 9472' if cfg == 5 else ''}"
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    vals = dict(zip(rows[0], rows[2])) if len(rows) > 2 else {}
    units = dict(zip(rows[0], rows[1])) if len(rows) > 2 else {}
    with open(os.path.join(P, f"{tag}_ncu_full_config{cfg}.txt"), "w") as f:
        f.write(f"ncu --set full --clock-control none --import-source on -k regex:sq_optimize -s 1 -c 1  {cmd}\n")
        f.write(open(os.path.join(G, f"full_c{cfg}.log")).read() if os.path.exists(os.path.join(G, f"full_c{cfg}.log")) else "")
        f.write("\n-- selected lines of `ncu --page details` --\n")
        for line in det.splitlines():
            if any(k in line for k in KEEP) or line.strip().startswith("void odam"):
                f.write(line.rstrip() + "\n")
        f.write("\n-- selected raw metrics --\n")
        for m in METRICS:
            if m in vals:
                f.write(f"{m} = {vals[m]} {units.get(m, '')}\n")
        mix = subprocess.run([sys.executable, os.path.join(REPO, "tools", "ncu_opcode_mix.py"), rep, "24"],
                             capture_output=True, text=True).stdout
        f.write("\n-- executed warp-instruction mix by SASS opcode (tools/ncu_opcode_mix.py) --\n" + mix)
        src = subprocess.run([sys.executable, os.path.join(REPO, "tools", "ncu_source_summary.py"), rep, "25"],
                             capture_output=True, text=True).stdout
        f.write("\n-- hottest CUDA source lines (tools/ncu_source_summary.py) --\n" + src)
        cyc = os.path.join(G, f"cycles_c{cfg}.log")
        if os.path.exists(cyc):
            f.write("\n-- per-phase SM cycles per iteration per object, thread 0's view (tools/prof_run.py --cycles; a separate, "
                    "un-profiled run) --\n" + open(cyc).read())
    try:
        traffic[f"config{cfg}"] = float(vals["dram__bytes_read.sum"].replace(",", "")) * (1e6 if "Mbyte" in units.get("dram__bytes_read.sum", "") else 1e3 if "Kbyte" in units.get("dram__bytes_read.sum", "") else 1) \
            + float(vals["dram__bytes_write.sum"].replace(",", "")) * (1e6 if "Mbyte" in units.get("dram__bytes_write.sum", "") else 1e3 if "Kbyte" in units.get("dram__bytes_write.sum", "") else 1)
    except (KeyError, ValueError):
        pass
if traffic:
    old = {}
    tp = os.path.join(P, "traffic.json")
    if os.path.exists(tp):
        old = json.load(open(tp))
    old.update(traffic)
    old["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of ONE sq_optimize_kernel launch from `ncu --set full` "
                    "(tools/prof_run.py; the command of each capture heads profiles/<round>_ncu_full_config<k>.txt) -- bytes "
                    "per launch of that capture; bench.py reports the config-2 figure as roofline.traffic")
    json.dump(old, open(tp, "w"), indent=1)
lc = os.path.join(G, "launches_bench.csv")
if os.path.exists(lc):
    lines = [l for l in open(lc) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = {}
    for r in rows:
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        u = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        name = r["Kernel Name"].split("(")[0][:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1
    with open(os.path.join(P, f"{tag}_launches_bench.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv  python bench.py --steps 5 --warmup 3 --no-cpu --no-sweep --no-strong --no-call-site\n"
                "(cold-cache, serialised launches: compare SHARES, not absolutes.  sq_optimize_kernel = 3 warm-up + 5 timed\n"
                " device-resident steps + 4 host-buffer (e2e) steps; inside the timed region of a step the only other launch is\n"
                " torch's 256 MiB L2-flush fill, outside the event pair; everything else is scene generation before the timed\n"
                " region and the FFMA roofline probe after it)\n\n")
        f.write(f"{'kernel':92s} {'launches':>8s} {'total ms':>10s} {'share':>7s}\n")
        for name, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{name:92s} {n:8d} {ms:10.3f} {ms / tot:7.1%}\n")
    import shutil
    shutil.copy(lc, os.path.join(P, f"{tag}_launches_bench.csv"))
for name in ("bench.json", "bench_ref.json", "r02_parity.txt", "callsite.log"):
    src = os.path.join(G, name)
    if os.path.exists(src) and os.path.getsize(src) > 0:
        with open(os.path.join(P, name if name.startswith(tag) else f"{tag}_{name}"), "w") as f:
            f.write(open(src).read())
print("profiles/:", sorted(os.listdir(P)))
