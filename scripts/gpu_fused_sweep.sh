#!/bin/bash
# sweep of the fused kernel's variants (resident CTAs per SM).  Usage (under gpurun): bash scripts/gpu_fused_sweep.sh TAG "4 3 5" [bench args] [run_tests]
set -u
TAG=${1:-sweep}; OUT=gpurun_out/$TAG; mkdir -p $OUT
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    sep = d.get("separate_kernels", {})
    print(round(d["value"] / 1e6, 1), "M reads/s", round(d["ms_per_step"], 4), "ms", {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items() if v},
          "| one pipeline", round(d.get("value_one_pipeline", {}).get("value", 0) / 1e6, 1), round(d.get("value_one_pipeline", {}).get("ms_per_step", 0), 4),
          "| e2e", round(d["e2e"]["value"] / 1e6, 1), "| roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3),
          "probes/s", round(d["roofline"]["probes_per_s"] / 1e9, 2), "G | separate", round(sep.get("ms_per_step", 0), 4))
except Exception as e:
    print("unreadable:", e)
PY
}
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
VARIANTS=${2:-4 3 5}
for c in $VARIANTS; do
  echo "== fused_ctas=$c ${3:-}"
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --opt fused_ctas=$c ${3:-} > $OUT/bench_$c.json 2> $OUT/bench_$c.err
  tail -1 $OUT/bench_$c.err | cut -c1-200; show $OUT/bench_$c.json
done
if [ -n "${4:-}" ]; then
  echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
fi
