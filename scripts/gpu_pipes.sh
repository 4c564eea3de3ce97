#!/bin/bash
# e2e / value against the number of pipelines kept in flight by the one host thread
OUT=gpurun_out/r02p; mkdir -p $OUT
for P in 2 3 4 6 8; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --pipelines $P > $OUT/p$P.json 2> $OUT/p$P.err
  python - <<PY
import json
d=json.loads([l for l in open("$OUT/p$P.json") if l.startswith("{")][-1])
print("pipes $P value %.3g (%.4f ms) e2e %.3g (%.3f ms) floor %.3f one-pipe %.3g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["pcie"]["floor_ms_per_step"], d["value_one_pipeline"]["value"]))
PY
done
