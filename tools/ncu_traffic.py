"""profiles/r2_traffic.json <- measured DRAM bytes per launch of the profiled kernels, from ncu `--set full` captures of bench.py.

python tools/ncu_traffic.py rays_per_launch rep1.ncu-rep [rep2.ncu-rep ...]
Per kernel: mean of dram__bytes_read.sum + dram__bytes_write.sum over the captured launches; render kernels also per ray
(bench.py scales `dram_bytes_per_ray` to its launch size for roofline.traffic; the others are per launch of the benched workload)."""
import csv, io, json, os, subprocess, sys
rpl = int(sys.argv[1])
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for rep in sys.argv[2:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    ig, it = hdr.index("Grid Size"), hdr.index("gpu__time_duration.sum")
    for r in rows[2:]:
        name = r[ik].split("(")[0].replace("void ", "").replace("pgrf::", "").split("<")[0].strip()
        b = float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
        e = out.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "rep": os.path.basename(rep), "grids": set(), "us": 0.0})
        e["launches"] += 1
        e["dram_bytes"] += b
        e["grids"].add(r[ig])
        e["us"] += float(r[it].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[it], 1e-3)
res = {}
for name, e in out.items():
    per = e["dram_bytes"] / e["launches"]
    res[name] = {"dram_bytes_per_launch": per, "launches": e["launches"], "grids": sorted(e["grids"]),
                 "us_per_launch_under_ncu": e["us"] / e["launches"],
                 "src": e["rep"] + " (ncu --set full of bench.py, dram__bytes_read.sum + dram__bytes_write.sum, mean over "
                 + str(e["launches"]) + " launches)"}
    if name.startswith("render_"):
        res[name]["dram_bytes_per_ray"] = per / rpl
        res[name]["rays_per_launch"] = rpl
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_traffic.json")
json.dump(res, open(dst, "w"), indent=1)
print(json.dumps(res, indent=1))
