"""Hot spots of one kernel from `ncu -i REP --page source --csv --kernel-name regex:PAT` output.
usage: python scripts/ncu_hot.py SRC.csv [top_n]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hi]
n = len(H)
data = [r for r in rows[hi + 1:] if len(r) >= n - 5 and r[0].startswith("0x")]
si, ii, src = H.index("# Samples"), H.index("Instructions Executed"), H.index("Source")
num = lambda x: int(x) if x.isdigit() else 0
print("kernel", rows[0][1][:80])
print("samples", sum(num(r[si]) for r in data), "sass lines", len(data), "warp instr executed", sum(num(r[ii]) for r in data))
hist, samp = collections.Counter(), collections.Counter()
for r in data:
    t = r[src].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    hist[op] += num(r[ii]); samp[op] += num(r[si])
print("opcode: executed, samples")
for op, c in hist.most_common(18):
    print(f"  {op:10s} {c:12d} {samp[op]:8d}")
cols = [c for c in H if c.startswith("stall_") and "Not Issued" not in c]
ci = [H.index(c) for c in cols]
tot = collections.Counter()
for r in data:
    for c, i in zip(cols, ci):
        tot[c] += num(r[i])
print("stall totals:", {k: v for k, v in tot.most_common(8)})
top = sorted(data, key=lambda r: -num(r[si]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 15]
for r in top:
    print(r[si], r[ii], r[src].strip()[:64], {c[6:]: r[i] for c, i in zip(cols, ci) if r[i] not in ("0", "")})
