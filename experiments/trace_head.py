"""Timeline of head_ts CTA 0 (UKBB_HEAD_DBG=16): cycles relative to the producer's first tile."""
import sys, numpy as np
t = np.fromfile(sys.argv[1], dtype=np.int64).reshape(12, 64)
names = ["P issue", "S0 issue", "E0 start", "E0 done", "S1 ready", "S1 issued", "E1 start", "E1 done", "S2 ready", "E2 start", "E2 done"]
t0 = t[0, 0]
print("tile " + " ".join("%9s" % n.replace(" ", "_") for n in names))
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print("%4d " % i + " ".join("%9d" % (t[e, i] - t0) if t[e, i] else "%9s" % "-" for e in range(11)))
d = np.diff(t[:11, 8:56], axis=1)
print("mean cycles per tile (tiles 8..55):", {names[e]: float(d[e].mean()) for e in range(11)})
lat = t[:11, 8:56] - t[1, 8:56]
print("mean latency from S0 issue:", {names[e]: float(lat[e].mean()) for e in range(11)})
