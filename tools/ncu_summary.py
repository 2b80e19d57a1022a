"""Key metrics + top stall lines of an .ncu-rep (run here, no GPU): python tools/ncu_summary.py file.ncu-rep [nlines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__average_warp_latency_issue_stalled", "smsp__warp_issue_stalled",
        "local_load", "local_store", "l1tex__data_bank_conflicts", "smsp__pcsamp_warps_issue_stalled"]
for r in rows[2:]:
    print("=" * 100)
    for h, u, v in zip(hdr, units, r):
        if any(k in h for k in keys) and v not in ("", "0"):
            if "pcsamp" in h or "issue_stalled" in h:
                try:
                    if float(v.replace(",", "")) < 1: continue
                except ValueError: pass
            print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = list(csv.reader(io.StringIO(src)))
# find header row
hi = next(i for i, l in enumerate(lines) if l and l[0] == "Address")
h = lines[hi]
ci = {n: i for i, n in enumerate(h)}
samp = ci.get("# Samples") or ci.get("Warp Stall Sampling (All Samples)")
body = [l for l in lines[hi + 1:] if len(l) == len(h)]
tot = sum(int(l[samp] or 0) for l in body)
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
print(f"--- top SASS lines by samples (total {tot})")
for l in sorted(body, key=lambda l: -int(l[samp] or 0))[:nl]:
    st = sorted(((int(l[ci[n]] or 0), n) for n in stall_cols), reverse=True)[:2]
    print(f"{int(l[samp] or 0):6d} {100.0 * int(l[samp] or 0) / max(tot, 1):5.1f}%  {l[ci['Source']][:90]:90s} {st}")
agg = {n: sum(int(l[ci[n]] or 0) for l in body) for n in stall_cols}
print("--- stall totals:", sorted(((v, k) for k, v in agg.items()), reverse=True)[:8])
