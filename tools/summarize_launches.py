"""Sums an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: python tools/summarize_launches.py in.csv out.csv "comment" """
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); im = hdr.index("Metric Name")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = r[ik].split("(")[0].replace("void ", "")
    base = name.split("<")[0]
    tot[base][0] += 1; tot[base][1] += float(r[iv].replace(",", "")) / 1e6
total = sum(v[1] for v in tot.values()); n = sum(v[0] for v in tot.values())
with open(sys.argv[2], "w") as f:
    f.write(f"# {sys.argv[3]}\n# total {total:.1f} ms over {n} launches\nkernel,launches,total_ms,share\n")
    for k, (c, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{c},{ms:.3f},{ms / total:.3f}\n")
print(open(sys.argv[2]).read())
