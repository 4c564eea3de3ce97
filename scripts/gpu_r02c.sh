#!/bin/bash
# r02c: full bench (with psikt drop-in, other configs, PCIe ceiling) + ncu launch list + ncu --set full of the fused kernel.
TAG=${1:-r02c}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c 'import __graft_entry__ as g; g.build()' > $OUT/build.log 2>&1 || tail -5 $OUT/build.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; tail -3 $OUT/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > $OUT/bench_under_ncu.json 2> $OUT/ncu_launches.err
tail -2 $OUT/ncu_launches.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:seeds_fused_kernel -s 8 -c 3 -o $OUT/prof \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --pipelines 1 > /dev/null 2> $OUT/ncu_full.err
tail -2 $OUT/ncu_full.err
ls -la $OUT
