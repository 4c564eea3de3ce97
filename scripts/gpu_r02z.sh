#!/bin/bash
# r02z: final pass on the final code -- full GPU suite, bench, ncu launch list, ncu --set full of the fused kernel and of
# the distance kernels, compute-sanitizer over the kernels added since r02s.
TAG=${1:-r02z}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; tail -2 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -1 $OUT/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > $OUT/bench_under_ncu.json 2> $OUT/ncu_launches.err
tail -1 $OUT/ncu_launches.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:seeds_fused_kernel -s 8 -c 3 -o $OUT/prof \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --pipelines 1 > /dev/null 2> $OUT/ncu_full.err
tail -1 $OUT/ncu_full.err
if [ "$2" = "dist" ]; then
cat > /tmp/dist_prof.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench
from psi_b200 import capi
g = bench.build_graph("chr22")
print(bench.distance_bench(torch, torch.device("cuda", 0), capi, g, 300, 500, n_pairs=2_000_000))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dist_ -c 8 -o $OUT/prof_dist python /tmp/dist_prof.py > $OUT/dist_prof.log 2> $OUT/ncu_dist.err
tail -1 $OUT/ncu_dist.err
fi
bash scripts/gpu_sanitizer2.sh $TAG
ls -la $OUT
