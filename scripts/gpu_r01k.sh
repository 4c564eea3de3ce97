#!/bin/bash
# r01k: direct ASCII seeding + resolve variants.  Usage (under gpurun): bash scripts/gpu_r01k.sh TAG
set -u
TAG=${1:-r01k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 | tee $OUT/smoke.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(round(d["value"] / 1e6, 1), "M reads/s", round(d["ms_per_step"], 4), "ms", {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items()},
          "e2e", round(d["e2e"]["value"] / 1e6, 1), "roofline", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("unreadable:", e)
PY
}
echo "== bench (defaults)" ; timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; tail -3 $OUT/bench.err ; show $OUT/bench.json
for v in "seeding_mode=1" "resolve_items=4" "resolve_ctas=5"; do
  tagv=$(echo "$v" | tr ' =' '__')
  opts=""; for o in $v; do opts="$opts --opt $o"; done
  echo "== bench $v"
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --pipelines 1 $opts > $OUT/bench_$tagv.json 2> $OUT/bench_$tagv.err
  show $OUT/bench_$tagv.json
done
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err ; cat $OUT/bench_ref.json
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipelines 1 > $OUT/bench_under_ncu.json 2> $OUT/ncu_launches.err
tail -2 $OUT/ncu_launches.err
echo "== ncu full (seed_reads, seeds_on_paths, compact_resolve)"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:seed_reads|seeds_on_paths|compact_resolve" -s 9 -c 6 -o $OUT/prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipelines 1 > /dev/null 2> $OUT/ncu_full.err
tail -2 $OUT/ncu_full.err
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/kernels_ncu_raw.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page details > $OUT/kernels_ncu_details.txt 2>/dev/null
ls -la $OUT
