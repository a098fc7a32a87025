"""Static code size per CUDA source line from an ncu report: how many SASS instructions each line compiled to
(all inlined copies), beside its share of executed instructions.  Finds what bloats the instruction cache.

    python tools/ncu_static_size.py gpurun_out/prof.ncu-rep [top_n]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
cur, key, hdr = None, None, None
static, execd, src, addrs = collections.Counter(), collections.Counter(), {}, collections.defaultdict(list)
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = ["Line No", "Source", "Address", "Sass"] + r[4:]
        continue
    if r[0] == "Function Name" or hdr is None or len(r) < 8:
        continue
    if r[2] == "-":
        key = (cur, int(r[0]))
        src[key] = r[1].strip()[:70]
        d = dict(zip(hdr, r))
        execd[key] = float(d.get("Instructions Executed") or 0)
    elif r[2].startswith("0x") and key is not None:
        static[key] += 1
        addrs[key].append(int(r[2], 16))
tot, te = sum(static.values()), sum(execd.values()) or 1
print(f"static SASS instructions {tot} ({tot * 16 / 1024:.0f} KB)")
byfile = collections.Counter()
for (f, _), n in static.items():
    byfile[f] += n
print("  " + "  ".join(f"{f}:{n}" for f, n in byfile.most_common()))
for (f, l), n in static.most_common(top):
    a = sorted(addrs[(f, l)])
    copies = 1 + sum(1 for x, y in zip(a, a[1:]) if y - x > 4096)
    print(f"{f[:14]:14s}:{l:4d} static {n:5d} (~{copies} sites) exec {execd[(f, l)] / te:6.2%}  {src.get((f, l), '')}")
