#!/bin/bash
# r02k: the default bench at N = 8, twice, after the clock-sampler fix.
OUT=gpurun_out/r02k; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for R in 1 2; do
  timeout 300 $TR --nproc-per-node $NG --master-port 2954$R bench.py --gpus $NG --steps 20 --warmup 5 > $OUT/bench_n${NG}_run$R.json 2> $OUT/bench_n${NG}_run$R.err
  python - <<PY
import json
d = json.loads(open("$OUT/bench_n${NG}_run$R.json").read().strip().splitlines()[-1])
print("run $R value %.4g ms/step %.4f e2e %.4g (%.3f ms) clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"]))
PY
done
