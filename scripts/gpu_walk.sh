timeout 1000 python -m pytest tests -m gpu -x -q -k "walk or fuzz or phases" 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("  value", round(d["value"] / 1e6, 1), round(d["ms_per_step"], 4), "one", round(d["value_one_pipeline"]["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items() if v})
PY
}
echo "== mhc walk 200k"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --shape mhc --k 32 --read-len 150 --reads 200000 --offpath-mode 1 > gpurun_out/r01zk_mhc_walk.json 2>/dev/null; show gpurun_out/r01zk_mhc_walk.json
echo "== mhc walk 1M"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --shape mhc --k 32 --read-len 150 --offpath-mode 1 > gpurun_out/r01zk_mhc_walk1m.json 2>/dev/null; show gpurun_out/r01zk_mhc_walk1m.json
echo "== chr22 walk 1M"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --offpath-mode 1 > gpurun_out/r01zk_chr22_walk.json 2>/dev/null; show gpurun_out/r01zk_chr22_walk.json
