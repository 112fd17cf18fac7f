"""Summarise an ncu CSV of ONE train step (tools/one_step.py) with the metrics
   gpu__time_duration.sum, sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed, dram__bytes_read.sum, dram__bytes_write.sum
into (a) a per-kernel table, (b) the time-weighted tensor-pipe % of the step and of the T5/ViT GEMM + attention kernels,
(c) DRAM bytes per launch of the dominant GEMM kernel.   python tools/summarize_step_profile.py in.csv out_prefix"""
import collections, csv, json, re, sys


def main(path, prefix):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        kid = row["ID"]
        d = per.setdefault(kid, {"name": re.sub(r"\(.*", "", row["Kernel Name"])[:90], "grid": row.get("Grid Size", "")})
        val = float(row["Metric Value"].replace(",", "")) if row["Metric Value"] not in ("", "n/a") else 0.0
        unit, name = row["Metric Unit"], row["Metric Name"]
        if name == "gpu__time_duration.sum":
            d["us"] = val / 1e3 if unit in ("ns", "nsecond") else (val * 1e3 if unit in ("ms", "msecond") else val)
        elif name.startswith("sm__pipe_tensor"):
            d["tensor_pct"] = val
        elif name.startswith("dram__bytes"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d["dram"] = d.get("dram", 0.0) + val * mult
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in per.values():
        a = agg[d["name"]]
        a[0] += 1; a[1] += d.get("us", 0.0); a[2] += d.get("us", 0.0) * d.get("tensor_pct", 0.0); a[3] += d.get("dram", 0.0)
    total = sum(a[1] for a in agg.values())
    out = [f"# {path}: {len(per)} launches of one train step, {total / 1e3:.2f} ms serialised (cold-cache under ncu: compare SHARES)",
           f"{'us':>10} {'launches':>8} {'share':>7} {'tensor%':>8} {'DRAM MB/launch':>15}  kernel"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{a[1]:10.1f} {a[0]:8d} {100 * a[1] / total:6.1f}% {a[2] / max(a[1], 1e-9):8.1f} {a[3] / a[0] / 1e6:15.1f}  {k}")
    open(prefix + "_launches.txt", "w").write("\n".join(out) + "\n")
    print("\n".join(out[:24]))
    mm = lambda n: ("gemm" in n) or ("attn_fwd" in n) or ("attn_bwd" in n)
    t_mm = sum(a[1] for k, a in agg.items() if mm(k))
    w_mm = sum(a[2] for k, a in agg.items() if mm(k))
    w_all = sum(a[2] for a in agg.values())
    gemm = {k: a for k, a in agg.items() if "gemm" in k}
    top = max(gemm.items(), key=lambda kv: kv[1][1])
    json.dump({"source": f"ncu launch list of one eager train step ({path}), metric sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed, "
                         "time-weighted over kernel durations (for tensor kernels time ~ FLOPs / rate, i.e. the FLOP-weighted figure)",
               "step_all_kernels_pct": w_all / total, "gemm_and_attention_kernels_pct": w_mm / max(t_mm, 1e-9),
               "gemm_kernels_pct": sum(a[2] for a in gemm.values()) / max(sum(a[1] for a in gemm.values()), 1e-9),
               "attention_kernels_pct": sum(a[2] for k, a in agg.items() if "attn_" in k and "delta" not in k) /
                                        max(sum(a[1] for k, a in agg.items() if "attn_" in k and "delta" not in k), 1e-9),
               "serialised_ms": total / 1e3}, open(prefix + "_tensor_pipe.json", "w"), indent=1)
    json.dump({"source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the {top[1][0]} launches of {top[0]} in one train step ({path})",
               "kernel": top[0], "dram_bytes_per_launch": top[1][3] / top[1][0]}, open(prefix + "_gemm_traffic.json", "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
