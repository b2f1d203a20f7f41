"""Extract per-launch DRAM traffic from .ncu-rep files into profiles/r02_ncu_traffic.json (TF21_TRAFFIC_JSON overrides).
usage: python tools/ncu_traffic.py shape_key=report.ncu-rep [...]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.environ.get("TF21_TRAFFIC_JSON") or os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
db = json.load(open(out_path)) if os.path.exists(out_path) else {}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for arg in sys.argv[1:]:
    key, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    entry = {}
    for r in data:
        name = r[ik].split("(")[0].replace("void ", "").replace("tf21::", "")
        tot = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
        entry.setdefault(name, []).append(tot)
    db[key] = {k: sum(v) / len(v) for k, v in entry.items()}
json.dump(db, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(db, indent=1))
