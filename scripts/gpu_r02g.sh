#!/bin/bash
# r02g: sliced index build -- parity tests, a forced-slices run at 1/16 scale, then BASELINE configs[4] at full size on one GPU.
TAG=${1:-r02g}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "slices or distance or seed_finder_api or saved_path_index or groups_of_paths" > $OUT/pytest_sel.log 2>&1; tail -4 $OUT/pytest_sel.log
timeout 600 python bench.py --shape wg_1_16 --reads-total 2000000 --read-len 150 --reads 1000000 --opt build_slices=16 > $OUT/wg_1_16_sliced.json 2> $OUT/wg_1_16_sliced.err
echo "wg_1_16 sliced rc=$?"; tail -2 $OUT/wg_1_16_sliced.err; cut -c1-400 $OUT/wg_1_16_sliced.json
timeout 600 python bench.py --shape wg_1_16 --reads-total 2000000 --read-len 150 --reads 1000000 > $OUT/wg_1_16_oneshot.json 2> $OUT/wg_1_16_oneshot.err
echo "wg_1_16 one-shot rc=$?"; tail -1 $OUT/wg_1_16_oneshot.err
if grep -q '"verified"' $OUT/wg_1_16_sliced.json; then
  (while true; do nvidia-smi --query-gpu=memory.used --format=csv,noheader; free -g | sed -n 2p; sleep 10; done) > $OUT/wg_mem.log 2>&1 &
  MON=$!
  timeout 2400 python bench.py --shape wg --reads-total 12500000 --read-len 150 --reads 1250000 > $OUT/wg_full_n1.json 2> $OUT/wg_full_n1.err
  echo "wg full rc=$?"; tail -4 $OUT/wg_full_n1.err; cut -c1-1200 $OUT/wg_full_n1.json
  kill $MON
  sort -n $OUT/wg_mem.log | tail -1
fi
