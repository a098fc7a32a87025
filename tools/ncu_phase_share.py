import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur_file, hdr, data = None, None, []
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = ["Line No", "Source", "Address", "Sass"] + r[4:]; continue
    if r[0] == "Function Name" or hdr is None or len(r) < 8 or r[2] != "-": continue
    d = dict(zip(hdr, r))
    try: d["_inst"] = float(d["Instructions Executed"] or 0); d["_samp"] = float(d["# Samples"] or 0)
    except: continue
    d["_file"] = cur_file; d["_line"]=int(d["Line No"]); data.append(d)
dev = [(82,109,"B0 ratio/count math"),(113,125,"node_powers"),(128,139,"F slot_logs"),(147,177,"B0.1 pool_powers"),(181,197,"B0.2 pool_ratios"),(202,261,"B0.3 pool_place"),(267,349,"B walk"),(351,378,"div helpers"),(382,435,"C cdf"),(438,451,"D lower_bound"),(453,483,"D/F point math"),(485,526,"cluster plumbing"),(528,557,"E project_uv/minmax"),(562,569,"E view_all_valid")]
ker = [(82,103,"G derive_param"),(108,179,"sampler driver+barriers"),(181,214,"D points"),(218,271,"E scan"),(275,302,"F resolve_arg"),(304,397,"prologue"),(398,421,"E driver"),(423,506,"F body"),(508,545,"G reduce"),(546,617,"G adam"),(618,640,"epilogue")]
agg = {}
ti = sum(d["_inst"] for d in data); ts=sum(d["_samp"] for d in data)
for d in data:
    name = d["_file"]
    tabl = dev if d["_file"]=="sq_device.cuh" else ker if d["_file"]=="sq_kernels.cu" else None
    if tabl:
        for lo,hi,n in tabl:
            if lo<=d["_line"]<=hi: name=n; break
        else: name = d["_file"]+":other"
    a = agg.setdefault(name,[0,0]); a[0]+=d["_inst"]; a[1]+=d["_samp"]
for n,(i,s) in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    print(f"{n:32s} inst {i/ti:6.2%}  samples {s/ts:6.2%}")
print("total inst", ti)
