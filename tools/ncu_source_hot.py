"""Top stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv`:  python tools/ncu_source_hot.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = rows[hdr_i + 1:]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
print(f"{len(data)} SASS instructions, {tot} samples; kernel: {rows[0][1][:100]}")
top = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for i in sorted(top):
    r = data[i]
    n = int(r[col["# Samples"]] or 0)
    why = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:3]
    prev = data[i - 1][col["Source"]].strip()[:50] if i else ""
    print(f"sass {i:5d} {100 * n / tot:5.1f}%  exec {r[col['Instructions Executed']]:>8}  {r[col['Source']].strip()[:60]:60s} | " +
          ", ".join(f"{w}:{c}" for c, w in why if c) + f"   [prev: {prev}]")
