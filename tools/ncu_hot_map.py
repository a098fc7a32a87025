"""Address map of the code an ncu capture actually executed: contiguous hot regions, their size, how often they ran
per (object, iteration) and which source lines they came from.  For instruction-cache work.

    python tools/ncu_hot_map.py gpurun_out/prof.ncu-rep OBJECTS ITERS
"""
import collections
import csv
import io
import subprocess
import sys

rep, n_obj, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
hdr, rows, cur, key = None, {}, None, None
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = ["Line No", "Source", "Address", "Sass"] + r[4:]
        continue
    if hdr is None or len(r) < 8:
        continue
    if r[2] == "-":
        key = f"{cur.replace('sq_', '').split('.')[0]}:{r[0]}"
        continue
    if r[2].startswith("0x"):
        d = dict(zip(hdr, r))
        rows[int(r[2], 16)] = (float(d["Instructions Executed"] or 0), key, float(d["# Samples"] or 0),
                               float(d.get("stall_no_inst", 0) or 0))
a = sorted(rows)
base, unit = a[0], n_obj * iters
hot = [x for x in a if rows[x][0] >= 0.5 * unit]
print(f"{len(a)} instructions, {len(hot)} executed at least once per 2 object-iterations ({len(hot) * 16 / 1024:.1f} KB)")
regions, start, prev = [], None, None
for x in hot:
    if start is None or x - prev > 16 * 12:   # a gap of > 12 cold instructions ends a region
        if start is not None:
            regions.append((start, prev))
        start = x
    prev = x
regions.append((start, prev))
tot_s = sum(v[2] for v in rows.values()) or 1
for s, e in regions:
    xs = [x for x in a if s <= x <= e]
    n = len(xs)
    if n < 8:
        continue
    per = sum(rows[x][0] for x in xs) / n / unit
    lines = collections.Counter(rows[x][1] for x in xs)
    smp = sum(rows[x][2] for x in xs)
    ni = sum(rows[x][3] for x in xs)
    print(f"{(s - base) / 1024:7.2f}KB +{n:4d} instr  x{per:6.2f}/obj-iter  samples {smp / tot_s:5.1%} (no_inst {ni / max(smp, 1):4.0%})  "
          + " ".join(f"{k}({v})" for k, v in lines.most_common(4)))
