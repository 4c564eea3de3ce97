#!/bin/bash
# r01n: compact (4 x u32) records + pipelined resident step.  Usage (under gpurun): bash scripts/gpu_r01n.sh TAG [skip_tests]
set -u
TAG=${1:-r01n}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
if [ -z "${2:-}" ]; then
  echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
  echo "== smoke" ; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 | tee $OUT/smoke.log
fi
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"] / 1e6, 1), "M reads/s", round(d["ms_per_step"], 4), "ms", {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items()},
          "| pipelined", round(d.get("value_pipelined", {}).get("value", 0) / 1e6, 1),
          "| e2e", round(d["e2e"]["value"] / 1e6, 1), round(d["e2e"]["ms_per_step"], 3), "ms",
          "| e2e wide", round(d.get("e2e_wide_records", {}).get("value", 0) / 1e6, 1), "| roofline", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("unreadable:", e)
PY
}
echo "== bench (defaults)" ; timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; tail -3 $OUT/bench.err ; show $OUT/bench.json
for p in 1 3; do
  echo "== bench --pipelines $p"
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --pipelines $p > $OUT/bench_p$p.json 2> $OUT/bench_p$p.err
  tail -1 $OUT/bench_p$p.err ; show $OUT/bench_p$p.json
done
ls -la $OUT
