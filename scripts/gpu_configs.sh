#!/bin/bash
# The other BASELINE configs as bench lines: chr22 shape with 150 bp reads (configs[2] read shape) and the MHC shape at
# k = 32 (configs[3]).  Usage (under gpurun): bash scripts/gpu_configs.sh TAG
set -u
TAG=${1:-cfg}; OUT=gpurun_out/$TAG; mkdir -p $OUT
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["workload"]); print(" ", d.get("route"), "|", d["offpath_mode"])
    print("  value", round(d["value"] / 1e6, 1), "M reads/s", round(d["ms_per_step"], 4), "ms; seeds/s", round(d["seeds_per_s"] / 1e9, 2), "G",
          {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items() if v}, "| e2e", round(d["e2e"]["value"] / 1e6, 1),
          "| roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), "probes/s", round(d["roofline"]["probes_per_s"] / 1e9, 2), "G",
          "| index", d["index"]["slot_bytes"], "B slots", round(d["index"]["bytes"] / 1e6), "MB", "slow/step", d["probe_slow_seeds_per_step"])
except Exception as e:
    print("unreadable:", e)
PY
}
echo "== chr22 shape, 150 bp reads, k = 20"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --read-len 150 > $OUT/bench_chr22_150.json 2> $OUT/bench_chr22_150.err; tail -1 $OUT/bench_chr22_150.err | cut -c1-250; show $OUT/bench_chr22_150.json
echo "== MHC shape, 150 bp reads, k = 32"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --shape mhc --k 32 --read-len 150 > $OUT/bench_mhc_k32.json 2> $OUT/bench_mhc_k32.err; tail -1 $OUT/bench_mhc_k32.err | cut -c1-250; show $OUT/bench_mhc_k32.json
echo "== MHC shape, walk mode"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --shape mhc --k 32 --read-len 150 --reads 200000 --offpath-mode 1 > $OUT/bench_mhc_k32_walk.json 2> $OUT/bench_mhc_k32_walk.err; tail -1 $OUT/bench_mhc_k32_walk.err | cut -c1-250; show $OUT/bench_mhc_k32_walk.json
