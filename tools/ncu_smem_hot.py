"""Shared-memory wavefronts per CUDA source line of an ncu report (total / excessive), per unit of work.
    python tools/ncu_smem_hot.py report.ncu-rep units [top]"""
import csv, io, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur, hdr, rows = None, None, []
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No":
        hdr = r; iW = hdr.index("L1 Wavefronts Shared"); iX = hdr.index("L1 Wavefronts Shared Excessive"); iE = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) <= iX or r[0] == "": continue
    try: rows.append((cur, int(r[0]), r[1].strip(), float(r[iW] or 0), float(r[iX] or 0), float(r[iE] or 0)))
    except ValueError: pass
tw, tx, ti = sum(x[3] for x in rows), sum(x[4] for x in rows), sum(x[5] for x in rows)
print(f"per unit: shared wavefronts {tw / units:.0f} (excessive {tx / units:.0f}), warp instructions {ti / units:.0f}")
for x in sorted(rows, key=lambda x: -x[3])[:top]:
    print(f"{x[0][:13]:13s}:{x[1]:4d} wavefronts {x[3] / units:7.1f} excess {x[4] / units:7.1f} inst {x[5] / units:6.1f}  {x[2][:90]}")
