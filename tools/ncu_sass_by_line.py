"""SASS instructions (with executed counts) attributed to a range of CUDA source lines of an ncu report.
    python tools/ncu_sass_by_line.py report.ncu-rep sq_kernels.cu 187 214 [units]
units: divide the counts by this number (e.g. objects x iterations) to get warp instructions per unit."""
import csv, io, subprocess, sys
rep, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
units = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur, line, src, hdr, tot = None, None, "", None, 0.0
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; iE = hdr.index("Instructions Executed"); iT = hdr.index("Avg. Threads Executed"); continue
    if hdr is None or len(r) <= iE: continue
    if r[0] != "": 
        try: line = int(r[0]); src = r[1]
        except ValueError: line = None
        if cur == fname and line is not None and lo <= line <= hi:
            print(f"--- {line}: {src.strip()[:100]}   [{float(r[iE] or 0) / units:.1f}]")
        continue
    if cur == fname and line is not None and lo <= line <= hi and r[3] not in ("...", ""):
        try: ie = float(r[iE])
        except ValueError: continue
        tot += ie
        print(f"      {r[3].strip()[:64]:64s} {ie / units:9.2f}  thr {r[iT]}")
print("total", tot / units)
