"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals (for profiles/)."""
import collections, csv, re, sys

def main(path, out=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        val = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        us = val / 1e3 if unit == "ns" else (val * 1e3 if unit == "ms" else val)
        key = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
        tot[key][0] += 1
        tot[key][1] += us
        n += 1
    s = sum(v[1] for v in tot.values())
    lines_out = [f"# {path}: {n} launches, {s/1e3:.2f} ms total (cold-cache, serialised: compare SHARES)",
                 f"{'us':>12} {'launches':>8} {'share':>7}  kernel"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        lines_out.append(f"{v[1]:12.1f} {v[0]:8d} {100*v[1]/s:6.1f}%  {k}")
    text = "\n".join(lines_out)
    print(text)
    if out:
        open(out, "w").write(text + "\n")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
