#!/bin/bash
# r02e: 8-GPU box: configs[4] at 1/4 scale on 8 GPUs, configs[2] on 8/4/2 GPUs, the regular bench at N = 8.
OUT=gpurun_out/r02e; mkdir -p $OUT
nvidia-smi -L | wc -l; free -g | head -2 | tail -1; nproc
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
MEM=$(free -g | awk 'NR==2{print $2}')
SHAPE=wg_1_4; TOTAL=25000000
if [ "$MEM" -lt 150 ]; then SHAPE=wg_1_16; TOTAL=6250000; fi
echo "config4 shape $SHAPE reads $TOTAL"
timeout 900 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --shape $SHAPE --reads-total $TOTAL --read-len 150 --reads 1600000 \
    > $OUT/config4_n8.json 2> $OUT/config4_n8.err; grep "sharded\]" $OUT/config4_n8.err | tail -1; tail -2 $OUT/config4_n8.err | cut -c1-300
for N in 8 4 2; do
  timeout 400 $TR --nproc-per-node $N --master-port 2952$N bench.py --gpus $N --reads-total 10000000 --read-len 150 --reads 2500000 \
      > $OUT/config2_n$N.json 2> $OUT/config2_n$N.err; tail -1 $OUT/config2_n$N.err | cut -c1-200
done
timeout 400 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs \
    > $OUT/bench_n8.json 2> $OUT/bench_n8.err; tail -1 $OUT/bench_n8.err | cut -c1-200
python - <<PY
import json
for f in ("config4_n8","config2_n8","config2_n4","config2_n2","bench_n8"):
    try:
        d=json.load(open("$OUT/%s.json"%f))
        print(f, "value %.4g e2e %.4g ms/step %.4f" % (d["value"] or 0, d["e2e"]["value"], d["ms_per_step"]), d.get("verified",""), d["index"]["bytes"])
    except Exception as e:
        print(f, "FAILED", e)
PY
