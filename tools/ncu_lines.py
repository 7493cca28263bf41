"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line:
warp instructions executed and stall samples, top lines first."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] in ("Function Name", "Kernel Name"): continue
    if hdr is None: continue
    if r[0] != "":   # source line row carries the totals for its SASS
        try:
            d = dict(zip(hdr[4:], r[4:]))
            inst = int(d["Instructions Executed"]); samp = int(d["# Samples"])
            tinst = int(d["Thread Instructions Executed"])
        except Exception: continue
        agg[(cur_file, int(r[0]))] = (inst, samp, tinst, r[1].strip()[:110], d)
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f"total warp-inst {tot_i:,}  samples {tot_s:,}")
for (f, ln), (inst, samp, tinst, src, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
    top3 = ",".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{f}:{ln:4d} inst {100*inst/tot_i:5.2f}% samp {100*samp/tot_s:5.2f}% act {tinst/max(inst,1):4.1f} | {top3} | {src}")
