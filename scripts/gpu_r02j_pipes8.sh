#!/bin/bash
# r02j: value / e2e at N = 8 against the number of pipelines per rank (the default became 6 after the last N = 8 run).
OUT=gpurun_out/r02j; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l); nproc
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for P in 4 6 3; do
  timeout 300 $TR --nproc-per-node $NG --master-port 2953$P bench.py --gpus $NG --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --pipelines $P \
      > $OUT/bench_n${NG}_p$P.json 2> $OUT/bench_n${NG}_p$P.err
  python - <<PY
import json
d = json.loads(open("$OUT/bench_n${NG}_p$P.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("pipelines $P value %.4g ms/step %.4f e2e %.4g (%.3f ms) launch_ms %.4f concurrency %.2f one-pipeline %.4f" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], r["launch_ms"], r["concurrency"], d["value_one_pipeline"]["ms_per_step"]))
PY
done
