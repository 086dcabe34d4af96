"""Per-source-line instruction / stall-sample totals of one kernel from an ncu report (needs -lineinfo).
usage: python scripts/ncu_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file, hdr = None, None
inst, samp, src = collections.Counter(), collections.Counter(), {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and len(r) > ie and r[0].isdigit() and r[2] == "-":      # a CUDA source line (aggregated over its SASS)
        key = (cur_file, int(r[0]))
        inst[key] += int(r[ie] or 0)
        samp[key] += int(r[isamp] or 0)
        src[key] = r[1].strip()
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {ti}  samples {ts}")
for k, v in inst.most_common(top):
    print(f"{k[0]}:{k[1]:5d} inst {100 * v / ti:5.1f}%  samples {100 * samp[k] / max(ts, 1):5.1f}%  {src[k][:110]}")
