#!/bin/bash
# Full pass of a round: parity tests, smoke, bench (with the CPU baseline), the other BASELINE configs, ncu launch list and
# full capture of the fused kernel.  Usage (under gpurun): bash scripts/gpu_final.sh TAG
set -u
TAG=${1:-r01zf}
bash scripts/gpu_r01t.sh $TAG
bash scripts/gpu_configs.sh $TAG
echo "== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/$TAG/bench_ref.json 2> gpurun_out/$TAG/bench_ref.err; cut -c1-300 gpurun_out/$TAG/bench_ref.json
