"""Summarise an ncu report per CUDA source line: share of executed warp instructions and of stall samples.

    python tools/ncu_source_summary.py gpurun_out/prof.ncu-rep [top_n]

Uses `ncu --page source --print-source cuda,sass --csv`; rows whose Address is "-" are the per-source-line
aggregates (compile with -lineinfo).
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
cur_file, hdr, data = None, None, []
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = ["Line No", "Source", "Address", "Sass"] + r[4:]
        continue
    if r[0] == "Function Name" or hdr is None or len(r) < 8 or r[2] != "-":
        continue
    d = dict(zip(hdr, r))
    try:
        d["_inst"] = float(d["Instructions Executed"] or 0)
        d["_samp"] = float(d["# Samples"] or 0)
        d["_thr"] = float(d["Avg. Threads Executed"] or 0)
    except (ValueError, KeyError):
        continue
    d["_file"] = cur_file
    data.append(d)
ti = sum(d["_inst"] for d in data) or 1
ts = sum(d["_samp"] for d in data) or 1
print(f"total warp instructions {ti:.4g}, stall samples {ts:.4g}")
by_file = {}
for d in data:
    a = by_file.setdefault(d["_file"], [0, 0])
    a[0] += d["_inst"]; a[1] += d["_samp"]
for f, (i, s) in by_file.items():
    print(f"  {f}: inst {i / ti:6.2%} samples {s / ts:6.2%}")
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
tot = {c: sum(float(d.get(c, 0) or 0) for d in data) for c in stall_cols}
print("stall reasons over the whole kernel: " +
      " ".join(f"{c[6:]}={v / ts:.1%}" for c, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v / ts >= 0.01))
for d in sorted(data, key=lambda d: -d["_samp"])[:top]:
    st = sorted(((float(d.get(c, 0) or 0), c) for c in stall_cols), reverse=True)[:2]
    sts = " ".join(f"{c[6:]}={v / max(d['_samp'], 1):.0%}" for v, c in st if v > 0)
    print(f"{d['_file'][:13]:13s}:{d['Line No']:>4s} inst {d['_inst'] / ti:6.2%} samp {d['_samp'] / ts:6.2%} "
          f"[{sts}] {d['Source'].strip()[:80]}")
