for sl in 0 1; do
echo "== chr22 index_slack=$sl"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --opt index_slack=$sl > gpurun_out/r02c_chr22_$sl.json 2> gpurun_out/r02c.err; python - gpurun_out/r02c_chr22_$sl.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("  value", round(d["value"] / 1e6, 1), round(d["ms_per_step"], 4), "one", round(d["value_one_pipeline"]["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items() if v}, "index MB", round(d["index"]["bytes"] / 1e6), "slow", d["probe_slow_seeds_per_step"], "sep probe", round(d["separate_kernels"]["kernel_ms_per_step"]["ms_probe"], 4), "build ms", round(d["index"]["build_ms"]))
PY
echo "== mhc index_slack=$sl"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --shape mhc --k 32 --read-len 150 --opt index_slack=$sl > gpurun_out/r02c_mhc_$sl.json 2> gpurun_out/r02c.err; python - gpurun_out/r02c_mhc_$sl.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("  value", round(d["value"] / 1e6, 1), round(d["ms_per_step"], 4), "one", round(d["value_one_pipeline"]["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items() if v}, "index MB", round(d["index"]["bytes"] / 1e6), "slow", d["probe_slow_seeds_per_step"], "sep probe", round(d["separate_kernels"]["kernel_ms_per_step"]["ms_probe"], 4), "build ms", round(d["index"]["build_ms"]))
PY
done
