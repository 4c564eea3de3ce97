#!/bin/bash
# r02h: BASELINE configs[4] at full size on one GPU (sliced index build), one GPU's share of the read set.
TAG=${1:-r02h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
(while true; do nvidia-smi --query-gpu=memory.used --format=csv,noheader; free -g | sed -n 2p; sleep 10; done) > $OUT/wg_mem.log 2>&1 &
MON=$!
timeout 2700 python bench.py --shape wg --reads-total 12500000 --read-len 150 --reads 1250000 > $OUT/wg_full_n1.json 2> $OUT/wg_full_n1.err
echo "wg full rc=$?"; tail -6 $OUT/wg_full_n1.err; cut -c1-1500 $OUT/wg_full_n1.json
kill $MON
grep MiB $OUT/wg_mem.log | sort -n | tail -1; grep Mem $OUT/wg_mem.log | sort -k3 -n | tail -1
