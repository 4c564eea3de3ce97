#!/bin/bash
# r02a: parity of the async test, first bench of the packed/dense/async path.  Usage: bash scripts/gpu_r02a.sh TAG
TAG=${1:-r02a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "async or dense_refuses or equal_length" 2>&1 | tail -15 | tee $OUT/pytest_async.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; tail -5 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value %.3g reads/s  ms/step %.4f | e2e %.3g (%.3f ms) | fused kernel %.4f ms  frac %.3f line-frac %.3f probes/s %.3g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["launch_ms"], d["roofline"]["frac"], d["roofline"]["line_accounting"]["frac"], d["roofline"]["probes_per_s"]))
for k in ("value_ascii_chunks","value_ascii_chunks_records","value_one_pipeline","e2e_ascii_chunks_records"):
    print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in d[k].items() if a!="note"})
print("separate", d["separate_kernels"]["kernel_ms_per_step"], d["separate_kernels"]["ms_per_step"])
print("other", json.dumps(d.get("other_configs"), indent=1)[:3000])
PY
