"""Top stall lines per kernel of an `ncu --page source --csv --print-source sass` dump.
usage: src_top.py dump.csv [top] [kernel-substring] [occurrence]"""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else ""
occ = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rows = list(csv.reader(open(path)))
secs = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}; secs.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
sel = [s for s in secs if want in s["name"]]
print("kernels:", len(secs), "matching:", len(sel))
s = sel[occ]
print(s["name"])
hdr, data = s["hdr"], s["data"]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top]
for i in sorted(order):
    r = data[i]
    st = {h[6:]: int(r[ix[h]]) for h in stalls if int(r[ix[h]]) > 0}
    top3 = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5d %6s %5.1f%% exec=%-8s %-72s %s" % (i, r[ix["# Samples"]], 100.0 * int(r[ix["# Samples"]]) / max(tot, 1),
                                            r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:72], top3))
