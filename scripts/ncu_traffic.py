"""dram__bytes_read.sum + dram__bytes_write.sum per kernel from an .ncu-rep -> JSON (bench.py's roofline.traffic).
Usage: ncu_traffic.py report.ncu-rep out.json"""
import csv, io, json, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = {}
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("is::", "")
    b = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        b += float(r[col[k]].replace(",", "")) * scale.get(units[col[k]], 1)
    t = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * {"us": 1e-3, "ns": 1e-6, "ms": 1.0}.get(units[col["gpu__time_duration.sum"]], 1e-3)
    a = res.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "ms": 0.0})
    a["launches"] += 1; a["dram_bytes"] += b; a["ms"] += t
for a in res.values():
    a["dram_bytes_per_launch"] = a["dram_bytes"] / a["launches"]
json.dump(res, open(sys.argv[2], "w"), indent=1)
print(json.dumps(res, indent=1))
