#!/bin/bash
# r02m: DENSE5 parity + bench
OUT=gpurun_out/r02m; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "dense or async or smoke or fullsize" > $OUT/pytest_sel.log 2>&1; tail -3 $OUT/pytest_sel.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-other-configs > $OUT/bench.json 2> $OUT/bench.err; tail -2 $OUT/bench.err
python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value %.4g (%.4f ms) e2e %.4g (%.4f ms) floor %.3f; 6-byte e2e %.4g (%.4f ms); kernel alone %.4f" % (d["value"], d["ms_per_step"], e["value"], e["ms_per_step"],
      e["pcie"]["floor_ms_per_step"], e["with_6_byte_results"]["value"], e["with_6_byte_results"]["ms_per_step"], d["roofline"]["launch_ms_alone"]))
print(e["h2d_bytes_per_step"], e["d2h_bytes_per_step"], d["roofline"]["frac"], d["roofline"]["line_accounting"]["frac"])
PY
