#!/bin/bash
# ncu pass: launch list + full capture of kernels matching REGEX. Usage: bash scripts/gpu_ncu.sh TAG REGEX [bench args]
TAG=$1; RE=$2; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > $OUT/bench.json 2> $OUT/bench.err; tail -1 $OUT/bench.err
python -c "
import json; d=json.load(open('$OUT/bench.json')); print(d['kernel_ms_per_step'], d.get('probe_slow_seeds_per_step'), d['index'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipelines 1 "$@" > $OUT/bench_under_ncu.json 2> $OUT/ncu_launches.err
python scripts/launch_summary.py $OUT/launches.csv --tail-from DeviceScanInit --nth 4 | tail -14
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RE -s 6 -c 3 -o $OUT/prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipelines 1 "$@" > /dev/null 2> $OUT/ncu_full.err
tail -1 $OUT/ncu_full.err
