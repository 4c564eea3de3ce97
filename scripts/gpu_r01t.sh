#!/bin/bash
# r01t: fused one-pass kernel (seeding + probe + records).  Usage (under gpurun): bash scripts/gpu_r01t.sh TAG [skip_tests] [skip_ncu]
set -u
TAG=${1:-r01t}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
echo "== smoke" ; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 | tee $OUT/smoke.log
if [ -z "${2:-}" ]; then
  echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
fi
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    sep = d.get("separate_kernels", {})
    print(d.get("route"), "|", round(d["value"] / 1e6, 1), "M reads/s", round(d["ms_per_step"], 4), "ms", {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items()},
          "| one pipeline", round(d.get("value_one_pipeline", {}).get("value", 0) / 1e6, 1), round(d.get("value_one_pipeline", {}).get("ms_per_step", 0), 4),
          "| e2e", round(d["e2e"]["value"] / 1e6, 1), round(d["e2e"]["ms_per_step"], 3), "ms",
          "| e2e wide", round(d.get("e2e_wide_records", {}).get("value", 0) / 1e6, 1),
          "| roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), "probes/s", round(d["roofline"]["probes_per_s"] / 1e9, 2), "G",
          "| separate", round(sep.get("ms_per_step", 0), 4), {k: round(v, 4) for k, v in sep.get("kernel_ms_per_step", {}).items()},
          "probe frac", round(sep.get("roofline_probe", {}).get("frac", 0), 3))
except Exception as e:
    print("unreadable:", e)
PY
}
echo "== bench (defaults)" ; timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; tail -3 $OUT/bench.err ; show $OUT/bench.json
if [ -z "${3:-}" ]; then
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipelines 1 > $OUT/bench_under_ncu.json 2> $OUT/ncu_launches.err
  python scripts/launch_summary.py $OUT/launches.csv 2>&1 | tail -25
  echo "== ncu full capture of the fused kernel"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:seeds_fused_kernel -s 4 -c 2 -o $OUT/prof_fused \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipelines 1 > /dev/null 2> $OUT/ncu_full.err
  tail -2 $OUT/ncu_full.err
fi
ls -la $OUT
