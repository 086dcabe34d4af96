"""Print per-launch durations (us) of our kernels from an `ncu --metrics gpu__time_duration.sum --csv` log.
usage: python scripts/ncu_times.py LOG [last_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
out = [(r[ki].replace("vipsy::", "").split("(")[0][-44:], float(r[vi].replace(",", "")) / 1e3)
       for r in rows[hdr + 1:] if len(r) > vi and ("vipsy" in r[ki])]
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(out)
print("  ".join(f"{k}={v:.1f}us" for k, v in out[-n:]))
