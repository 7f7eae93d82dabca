"""profiles/r2_traffic.json <- DRAM bytes of the profiled kernels from an ncu capture of tools/profile_target.py.

python tools/ncu_traffic.py gpurun_out/prof_r2x.ncu-rep [rays_per_launch=16384]
bench.py scales `dram_bytes_per_ray` to its launch size (roofline.traffic); the cost-volume entry is per launch (configs[0])."""
import csv, io, json, os, subprocess, sys
rep = sys.argv[1]
rpl = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in rows[2:]:
    name = r[ik].split("(")[0].replace("void ", "").split("<")[0].strip()
    b = float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
    e = out.setdefault(name, {"launches": 0, "dram_bytes": 0.0})
    e["launches"] += 1
    e["dram_bytes"] += b
res = {}
for name, e in out.items():
    per = e["dram_bytes"] / e["launches"]
    res[name] = {"dram_bytes_per_launch": per, "src": os.path.basename(rep) + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean over "
                 + str(e["launches"]) + " launches)"}
    if name.startswith("render_") or name.startswith("project_gather") or name.startswith("depth_guided"):
        res[name]["dram_bytes_per_ray"] = per / rpl
        res[name]["rays_per_launch"] = rpl
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_traffic.json")
json.dump(res, open(dst, "w"), indent=1)
print(json.dumps(res, indent=1))
