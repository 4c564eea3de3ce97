#!/bin/bash
# One GPU pass: parity tests, bench, ncu launch list, ncu full capture of the probe kernel.
# Usage (under gpurun): bash scripts/gpu_pass.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; tail -3 $OUT/bench.err ; cat $OUT/bench.json
echo "== bench (walk mode, 1 pipeline)" ; timeout 900 python bench.py --steps 10 --warmup 3 --offpath-mode 1 --pipelines 1 --no-cpu-baseline > $OUT/bench_walk.json 2> $OUT/bench_walk.err ; tail -2 $OUT/bench_walk.err; cat $OUT/bench_walk.json
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err ; cat $OUT/bench_ref.json
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipelines 1 > $OUT/bench_under_ncu.json 2> $OUT/ncu_launches.err
tail -2 $OUT/ncu_launches.err
echo "== ncu full (seeds_on_paths, compact_resolve)"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:seeds_on_paths|compact_resolve" -s 6 -c 4 -o $OUT/prof_on \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipelines 1 > /dev/null 2> $OUT/ncu_full.err
tail -2 $OUT/ncu_full.err
ls -la $OUT
