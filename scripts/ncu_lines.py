"""Per-source-line instruction counts and stall samples from an ncu report (--import-source on).
usage: python scripts/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, collections, io, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; per = collections.OrderedDict(); src = {}; key = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1]; continue
    if len(r) <= 8 or r[0] == "Line No": continue
    if r[0] != "":
        key = (cur, int(r[0])); src[key] = r[1]; per.setdefault(key, [0, 0])
    elif key is not None:
        try: per[key][0] += int(r[7]); per[key][1] += int(r[4])
        except ValueError: pass
tot = sum(v[0] for v in per.values()); samp = sum(v[1] for v in per.values())
print("total warp instructions", tot, "samples", samp)
for (f, l), (n, s) in sorted(per.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{f.split('/')[-1]}:{l:4d} {n:9d} {100*n/tot:5.1f}%  samp {100*s/max(samp,1):5.1f}%  {src[(f,l)].strip()[:100]}")
