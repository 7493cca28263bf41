"""Sum ncu per-source-line metrics over line ranges of qz_deflate.cu (cuda,sass export)."""
import csv, sys
path = sys.argv[1]
ranges = [tuple(map(int, a.split('-'))) for a in sys.argv[2:]]
rows = list(csv.reader(open(path)))
cur = None; hdr = None; tot_i = tot_s = 0; acc = {r: [0, 0] for r in ranges}; other = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] in ("", "Function Name", "Kernel Name"): continue
    try:
        d = dict(zip(hdr[4:], r[4:])); inst = int(d["Instructions Executed"]); samp = int(d["# Samples"]); ln = int(r[0])
    except Exception: continue
    tot_i += inst; tot_s += samp
    if cur == "qz_deflate.cu":
        for rg in ranges:
            if rg[0] <= ln <= rg[1]: acc[rg][0] += inst; acc[rg][1] += samp
    else:
        o = other.setdefault(cur, [0, 0]); o[0] += inst; o[1] += samp
for rg in ranges: print(f"lines {rg[0]}-{rg[1]}: inst {100*acc[rg][0]/tot_i:5.2f}%  samples {100*acc[rg][1]/tot_s:5.2f}%")
for f, o in sorted(other.items(), key=lambda kv: -kv[1][1]): print(f"{f}: inst {100*o[0]/tot_i:5.2f}%  samples {100*o[1]/tot_s:5.2f}%")
print("total inst", tot_i, "samples", tot_s)
