"""Executed warp-instruction mix by SASS opcode from an ncu report (source page, sass view)."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(r for r in rows if r and r[0] == "Address")
ii, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
mix, smp = collections.Counter(), collections.Counter()
for r in rows:
    if len(r) <= ii or not r[0].startswith("0x"):
        continue
    toks = r[isrc].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0]
    mix[op] += float(r[ii] or 0); smp[op] += float(r[ismp] or 0)
tot, ts = sum(mix.values()), sum(smp.values())
print(f"total executed warp instructions {tot:.4g}")
for op, v in mix.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{op:10s} {v / tot:7.2%}   stall samples {smp[op] / ts:7.2%}")
