#!/bin/bash
# r02i: 8-GPU box: a cheap check of the shared-host-graph path (1/16 scale), the regular bench at N = 8 (weak scaling,
# six pipelines) and BASELINE configs[4] as written: 3.1 Gbp graph, 100 M x 150 bp reads sharded over 8 GPUs.
OUT=gpurun_out/r02i; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l); free -g | head -2 | tail -1; nproc; echo "gpus $NG"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --shape wg_1_16 --reads-total 2000000 --read-len 150 --reads 500000 \
    > $OUT/wg_1_16_n2.json 2> $OUT/wg_1_16_n2.err; echo "wg_1_16 n2 rc=$?"; grep "sharded\]" $OUT/wg_1_16_n2.err | tail -1 | cut -c1-300
if ! grep -q '"verified"' $OUT/wg_1_16_n2.json; then tail -5 $OUT/wg_1_16_n2.err; echo "shared-graph path failed: stopping"; exit 1; fi
timeout 400 $TR --nproc-per-node $NG --master-port 29531 bench.py --gpus $NG --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs \
    > $OUT/bench_n$NG.json 2> $OUT/bench_n$NG.err; tail -1 $OUT/bench_n$NG.err | cut -c1-200
(while true; do free -g | sed -n 2p; sleep 10; done) > $OUT/host_mem.log 2>&1 &
MON=$!
TOTAL=$((12500000 * NG))
timeout 1500 $TR --nproc-per-node $NG --master-port 29511 bench.py --gpus $NG --shape wg --reads-total $TOTAL --read-len 150 --reads 1250000 \
    > $OUT/config4_full_n$NG.json 2> $OUT/config4_full_n$NG.err; echo "config4 full rc=$?"
kill $MON
grep "sharded\]" $OUT/config4_full_n$NG.err | tail -1; tail -2 $OUT/config4_full_n$NG.err | cut -c1-300
sort -k3 -n $OUT/host_mem.log | tail -1
python - <<PY
import json
for f in ("wg_1_16_n2", "bench_n$NG", "config4_full_n$NG"):
    try:
        d = json.loads(open("$OUT/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %.4g ms/step %.4f" % (d["value"] or 0, d["e2e"]["value"], d["ms_per_step"]), d.get("verified", ""))
    except Exception as e:
        print(f, "FAILED", e)
PY
