#!/bin/bash
# full-size parity tests + e2e pipeline-count sweep.  Usage (under gpurun): bash scripts/gpu_exp3.sh TAG
TAG=${1:-exp3}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== full-size parity" ; timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q --durations=8 2>&1 | tail -25 | tee $OUT/pytest_fullsize.log
for p in 1 2 3 4; do
  echo "== bench pipelines=$p"
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --pipelines $p > $OUT/bench_p$p.json 2> $OUT/bench_p$p.err
  python - <<PY
import json; d=json.load(open("$OUT/bench_p$p.json")); print(d["value"], d["ms_per_step"], d["kernel_ms_per_step"], d["e2e"], d["roofline"]["frac"])
PY
done
