#!/bin/bash
# a sweep of one bench option (optionally the gpu tests first: TESTS=1). Usage: bash scripts/gpu_sweep.sh TAG name v1 v2 ...
TAG=$1; NAME=$2; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c 'import __graft_entry__ as g; g.build()' > $OUT/build.log 2>&1 || tail -5 $OUT/build.log
if [ "${TESTS:-0}" = "1" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.log; fi
for v in "$@"; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --opt $NAME=$v ${BENCH_ARGS:-} > $OUT/bench_$v.json 2> $OUT/bench_$v.err || tail -3 $OUT/bench_$v.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$v.json"))
print("$NAME=$v value %.3g reads/s  ms/step %.3f  e2e %.3g (%.2f ms)  roof frac %.3f  probes/s %.3g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["probes_per_s"]), d["kernel_ms_per_step"])
PY
done
