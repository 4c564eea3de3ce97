#!/bin/bash
# last pass of the round: full GPU suite, smoke(), the default bench and the reference arm
TAG=${1:-r02final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -2 $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -1 $OUT/bench.err
timeout 400 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err
python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
r = json.loads(open("$OUT/bench_ref.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g frac %.3f psikt %.3g ref %.4g clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["drop_in_psikt"]["reads_per_s"], r["value"], d["clocks"]))
PY
