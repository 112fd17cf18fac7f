mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --csv --log-file gpurun_out/r02_decode_step.csv python tools/one_decode_step.py > /dev/null 2>&1
python - <<'PY'
import csv, collections, re
lines=[l for l in open('gpurun_out/r02_decode_step.csv') if not l.startswith('==')]
agg=collections.defaultdict(lambda:[0,0.0,0.0])
for r in csv.DictReader(lines):
    k=re.sub(r"\(.*","",r["Kernel Name"])[:60]+" grid"+r["Grid Size"]
    v=float(r["Metric Value"].replace(",",""))
    if r["Metric Name"]=="gpu__time_duration.sum":
        agg[k][0]+=1; agg[k][1]+=v/1e3 if r["Metric Unit"] in ("ns","nsecond") else v
    else:
        agg[k][2]+=v*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}.get(r["Metric Unit"],1)
tot=sum(a[1] for a in agg.values())
print(f"one decode step: {sum(a[0] for a in agg.values())} launches, {tot:.0f} us serialised")
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:16]:
    print(f"{a[1]:8.1f} us {a[0]:4d}x  {a[2]/1e6/a[0]:8.1f} MB read/launch  {k}")
PY
