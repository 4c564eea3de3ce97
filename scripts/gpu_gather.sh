#!/bin/bash
OUT=gpurun_out/$1; mkdir -p $OUT
./bench_support/gather_peak 1024 5000000 10 > $OUT/gather_1G_5M.jsonl 2>&1
./bench_support/gather_peak 1024 50000000 5 > $OUT/gather_1G_50M.jsonl 2>&1
python - <<PY
import json
for f in ["gather_1G_5M","gather_1G_50M"]:
    print(f)
    for l in open("$OUT/"+f+".jsonl"):
        try: d=json.loads(l)
        except Exception: print(l.strip()); continue
        if d["test"]=="random_gather": print("  %-24s %6.2f G/s  %7.1f GB/s useful  %.4f ms" % (d["variant"], d["Gprobes_per_s"], d["useful_GBps"], d["ms"]))
PY
