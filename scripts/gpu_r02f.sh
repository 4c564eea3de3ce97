#!/bin/bash
# r02f: full GPU suite, bench (with the distance-index lines), TMA gather4 microbenchmark.
TAG=${1:-r02f}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv > $OUT/box.txt; free -g >> $OUT/box.txt; nproc >> $OUT/box.txt; df -h /tmp /dev/shm >> $OUT/box.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
for box in 4 1; do
  timeout 120 bench_support/gather_peak 2048 5000000 8 gather4 $box >> $OUT/gather4.jsonl 2>> $OUT/gather4.err
  timeout 120 bench_support/gather_peak 2048 50000000 5 gather4 $box >> $OUT/gather4.jsonl 2>> $OUT/gather4.err
done
cat $OUT/gather4.jsonl; tail -3 $OUT/gather4.err
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; tail -3 $OUT/bench.err
python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "e2e")})
print(json.dumps(d.get("other_configs", {}).get("distance_index_300_500"), indent=1))
PY
