#!/bin/bash
# experiment: random-sector gather ceiling + L2 fetch granularity effect on the bench
OUT=gpurun_out/r01b; mkdir -p $OUT
./bench_support/gather_peak 2048 5000000 10 > $OUT/gather_5M.jsonl 2>&1
./bench_support/gather_peak 2048 50000000 5 > $OUT/gather_50M.jsonl 2>&1
cat $OUT/gather_5M.jsonl $OUT/gather_50M.jsonl
for g in 0 128 64 32; do
  echo "== bench L2_FETCH=$g"
  PSI_B200_L2_FETCH=$g timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_l2fetch_$g.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
done
