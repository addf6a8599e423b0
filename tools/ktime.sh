#!/bin/bash
# usage: tools/ktime.sh <tag> <kernel-regex>  -> per-kernel durations of one bench step (ncu timing pass, cold-cache, serialised)
tag=$1; rx=$2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$rx" -c 60 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_$tag.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(l for l in open("gpurun_out/launches_$tag.csv") if not l.startswith("=="))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
acc=collections.OrderedDict()
for r in rows[1:]:
    acc.setdefault(r[ki][:70],[]).append(float(r[vi].replace(",",""))/1e3)
for k,v in acc.items(): print(f"{k:72s} n={len(v):3d} avg={sum(v)/len(v):10.1f} us min={min(v):10.1f}")
PY
