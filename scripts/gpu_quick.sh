#!/bin/bash
# quick pass: gpu tests + bench (no ncu). Usage: bash scripts/gpu_quick.sh TAG [extra bench args]
TAG=${1:-q}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
python -c 'import __graft_entry__ as g; g.build()' > $OUT/build.log 2>&1 || tail -5 $OUT/build.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > $OUT/bench.json 2> $OUT/bench.err; tail -2 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value %.3g reads/s  ms/step %.3f  e2e %.3g (%.2f ms)  roof frac %.3f  probes/s %.3g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["probes_per_s"]))
print(d["kernel_ms_per_step"])
PY
