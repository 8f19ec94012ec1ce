"""profiles/r02_traffic.json from an ncu CSV (--page raw --csv or --csv with --metrics) that holds, per launch,
gpu__time_duration.sum, dram__bytes_read.sum and dram__bytes_write.sum of the forward-type convolution kernels of one step.
    python scripts/make_traffic.py gpurun_out/conv_traffic.csv "source description"
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, src = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
rows = list(csv.reader(l for l in open(path) if l.startswith('"') or l[:1].isdigit() or l.startswith("ID") or l.startswith(",")))
hdr = rows[0]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
fam = {"fwd": [], "wgrad": []}


def add(name, nbytes):
    if "k_conv_fwd_umma" in name or "k_conv_first_fwd" in name:
        fam["fwd"].append(nbytes)
    elif "k_conv_wgrad_umma" in name or "k_conv_first_wgrad" in name:
        fam["wgrad"].append(nbytes)


if "Metric Name" in hdr:          # long format (ncu --csv --metrics ...): one row per (launch, metric)
    ki, ii, mi, ui, vi = (hdr.index(c) for c in ("Kernel Name", "ID", "Metric Name", "Metric Unit", "Metric Value"))
    per = {}
    for r in rows[1:]:
        if r[mi] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            e = per.setdefault(r[ii], [r[ki], 0.0])
            e[1] += float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1)
    for name, b in per.values():
        add(name, b)
else:                             # wide format (ncu -i rep --page raw --csv): second row holds the units
    units, body = rows[1], rows[2:]
    kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    for r in body:
        add(r[kn], float(r[rd].replace(",", "")) * UNIT.get(units[rd], 1) + float(r[wr].replace(",", "")) * UNIT.get(units[wr], 1))
out = {"kernel": "k_conv_fwd_umma* / k_conv_first_fwd (forward + data-gradient convolution family)", "source": src,
       "launches_captured": len(fam["fwd"]), "dram_bytes_per_launch_avg": sum(fam["fwd"]) / max(1, len(fam["fwd"])),
       "wgrad_launches_captured": len(fam["wgrad"]),
       "wgrad_dram_bytes_per_launch_avg": sum(fam["wgrad"]) / max(1, len(fam["wgrad"]))}
with open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
