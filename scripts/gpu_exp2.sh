#!/bin/bash
OUT=gpurun_out/r01c; mkdir -p $OUT
python -c 'import __graft_entry__ as g; g.build()' > $OUT/build.log 2>&1 || tail -5 $OUT/build.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
./bench_support/gather_peak 2048 5000000 10 > $OUT/gather_5M.jsonl 2>&1
./bench_support/gather_peak 2048 50000000 5 > $OUT/gather_50M.jsonl 2>&1
cat $OUT/gather_5M.jsonl $OUT/gather_50M.jsonl | cut -c1-300
