"""DRAM traffic of the dominant kernel from an `ncu --set full` report -> profiles/r2_head_traffic.json (read by bench.py's roofline object).
usage: ncu -i gpurun_out/full_<mode>.ncu-rep --page raw --csv | python experiments/ncu_traffic.py <mode> <report name>"""
import csv, json, os, sys
mode, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def col(name):
    return [i for h, i in ix.items() if h.endswith(name)][0]
rd, wr, nm = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), ix["Kernel Name"]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
tot = None
for r in rows[2:]:
    if "head_ts_kernel" in r[nm]:
        tot = float(r[rd].replace(",", "")) * scale[units[rd]] + float(r[wr].replace(",", "")) * scale[units[wr]]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "r2_head_traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d[mode] = {"bytes_per_launch": tot, "source": "dram__bytes_read.sum + dram__bytes_write.sum of one head_ts_kernel launch (500 SA slices), ncu --set full, " + src}
json.dump(d, open(path, "w"), indent=1)
print(mode, tot)
