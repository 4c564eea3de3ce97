show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("  value", round(d["value"] / 1e6, 1), round(d["ms_per_step"], 4), "one", round(d["value_one_pipeline"]["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items() if v}, "seeds/s", round(d["seeds_per_s"]/1e9,2))
PY
}
timeout 600 python -m pytest tests -m gpu -x -q -k "ragged or golden or edge or fuzz" 2>&1 | tail -2
echo "== default"; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s_def.json 2>/dev/null; show gpurun_out/s_def.json
echo "== 200k reads"; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --reads 200000 > gpurun_out/s_200k.json 2>/dev/null; show gpurun_out/s_200k.json
echo "== 200k reads fused_ctas... R=256 reference: n/a"
