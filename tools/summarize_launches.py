"""Condense an `ncu --metrics gpu__time_duration.sum --csv` log into id,kernel,grid,block,duration_us and a per-kernel summary.
usage: summarize_launches.py <ncu.csv> <out.csv> <out_summary.txt> "<header comment>" """
import csv, re, sys
from collections import defaultdict
src, out_csv, out_sum, note = sys.argv[1:5]
rows = []
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*$", "", name).strip()
    us = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] in ("nsecond", "ns"): us /= 1e3
    elif r["Metric Unit"] in ("msecond", "ms"): us *= 1e3
    rows.append((int(r["ID"]), name, r["Grid Size"], r["Block Size"], us))
with open(out_csv, "w") as f:
    f.write(f"# {note}\n")
    w = csv.writer(f)
    w.writerow(["id", "kernel", "grid", "block", "duration_us"])
    for r in rows: w.writerow([r[0], r[1], r[2], r[3], f"{r[4]:.3f}"])
tot = sum(r[4] for r in rows)
agg = defaultdict(lambda: [0.0, 0])
for r in rows:
    k = re.sub(r"^void ", "void ", r[1])[:110]
    agg[k][0] += r[4]; agg[k][1] += 1
with open(out_sum, "w") as f:
    f.write(f"{note}\n{len(rows)} launches, {tot/1e3:.1f} ms of kernel time (cold-cache, serialised under ncu)\n")
    f.write("share   ms     launches  kernel\n")
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        f.write(f"{100*t/tot:5.1f}%  {t/1e3:7.3f}  {n:6d}  {k}\n")
print(f"{len(rows)} launches, {tot/1e3:.2f} ms")
