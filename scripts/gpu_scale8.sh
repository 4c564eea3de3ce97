#!/bin/bash
# N = 8 and N = 1 on one 8-GPU box, as the driver launches them.  Usage (under gpurun --gpus 8): bash scripts/gpu_scale8.sh TAG
set -u
TAG=${1:-r01zz}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
show() { python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("n_gpus", d["n_gpus"], "value", round(d["value"] / 1e6, 1), "M reads/s", round(d["ms_per_step"], 4), "ms/step; one pipeline",
          round(d["value_one_pipeline"]["value"] / 1e6, 1), "; e2e", round(d["e2e"]["value"] / 1e6, 1), "M reads/s", round(d["e2e"]["ms_per_step"], 3),
          "ms/step; roofline", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("unreadable:", e)
PY
}
for N in 8 1; do
  echo "== bench --gpus $N"
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) \
        bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
  fi
  show $OUT/bench_n$N.json; tail -2 $OUT/bench_n$N.err | cut -c1-300
done
