#!/bin/bash
# Multi-GPU pass: bench.py under torchrun exactly as the driver launches it.  Usage (under gpurun --gpus N): bash scripts/gpu_multi.sh N TAG
set -u
N=${1:-2}; TAG=${2:-r01m}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.csv 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== bench --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
tail -5 $OUT/bench_n$N.err
python - $OUT/bench_n$N.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("n_gpus", d["n_gpus"], "value", round(d["value"] / 1e6, 1), "M reads/s", round(d["ms_per_step"], 4), "ms/step; e2e",
          round(d["e2e"]["value"] / 1e6, 1), "M reads/s", round(d["e2e"]["ms_per_step"], 3), "ms/step; roofline", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("unreadable:", e)
PY
echo "== bench --impl reference --gpus $N (rank 0 only)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 1 --warmup 0 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err
cat $OUT/bench_ref_n$N.json | cut -c1-400
echo "== bench --gpus 1 on the same box"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err
python - $OUT/bench_n1.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("n_gpus", d["n_gpus"], "value", round(d["value"] / 1e6, 1), "M reads/s; e2e", round(d["e2e"]["value"] / 1e6, 1), "M reads/s")
except Exception as e:
    print("unreadable:", e)
PY
ls -la $OUT
