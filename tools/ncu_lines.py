"""Aggregate an ncu source-page CSV per source line: python tools/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import collections, csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
sec = hdr = None
agg = collections.defaultdict(lambda: [0, 0, "", collections.Counter()])
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        sec = r[1]; i += 2; hdr = rows[i]; i += 1
        stall_cols = [(j, h) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr and r and r[0].isdigit():
        try:
            line = int(r[0]); inst = int(r[hdr.index("Instructions Executed")] or 0); smp = int(r[hdr.index("# Samples")] or 0)
        except Exception:
            i += 1; continue
        a = agg[(sec.split("/")[-1], line)]; a[0] += inst; a[1] += smp; a[2] = r[1][:110]
        for j, h in stall_cols:
            try: a[3][h] += int(r[j] or 0)
            except Exception: pass
    i += 1
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
allst = collections.Counter()
for v in agg.values(): allst.update(v[3])
print("instructions", tot, "samples", ts)
print({k: round(100 * v / max(ts, 1), 1) for k, v in allst.most_common(10)})
byfile = collections.Counter()
for k, v in agg.items(): byfile[k[0]] += v[0]
print({k: round(100 * v / tot, 1) for k, v in byfile.most_common(6)})
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:4d} inst {100*v[0]/tot:5.1f}% smp {100*v[1]/max(ts,1):5.1f}%  {v[2]}")
