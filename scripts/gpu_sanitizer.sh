#!/bin/bash
# compute-sanitizer memcheck + racecheck + synccheck over the tiny config (x.gfa, 10 000 x 100 bp reads, k = 12) through the
# three device routes and the new formats.  Usage: bash scripts/gpu_sanitizer.sh TAG
TAG=${1:-r02s}; OUT=gpurun_out/$TAG; mkdir -p $OUT
cat > /tmp/san_case.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, util
from psi_b200 import capi
case = {c["name"]: c for c in util.golden_index()["cases"]}[sys.argv[1]]
g = capi.Graph.load_gfa(util.GOLDEN / case["gfa"])
rp, bases = util.read_fasta(util.GOLDEN / case["reads"])
n = min(len(rp) - 1, 2000); rp = rp[:n + 1]; bases = bases[:int(rp[-1])]
for mode, fused in ((0, 1), (0, 0), (1, 1)):
    ctx = capi.Context(case["k"], 0)
    ctx.set_option("offpath_mode", mode); ctx.set_option("fused", fused)
    ctx.set_graph(g, ids="coord"); ctx.set_paths(g.pick_paths(case["n_paths"], seed=1)); ctx.find_loci()
    ctx.submit_chunk(rp, bases, 0, case["d"]); a = ctx.seeds_all(); rec = capi.canonical(ctx.fetch())
    ctx.submit_chunk_packed(capi.Packed.pack(rp, bases, 0), case["d"]); b = ctx.seeds_all(capi.ALL | capi.COMPACT); ctx.fetch32()
    assert a == b == len(rec)
    if mode == 0:
        c = ctx.seeds_all(capi.ALL | capi.DENSE); d, e = ctx.fetch_dense()
        r2, _ = capi.dense_to_records(d, e, rp, case["k"], case["d"], 0)
        assert c == a and np.array_equal(capi.canonical(r2), rec)
        ctx.seeds_all(capi.ALL | capi.SORTED); ctx.fetch()
    ctx.close()
    print("route", mode, fused, "hits", a, flush=True)
PY
for tool in memcheck racecheck synccheck; do
  for c in x_k12 m_k32; do
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_case.py $c > $OUT/${tool}_$c.log 2>&1
    echo "$tool $c rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/${tool}_$c.log | tail -1)"
  done
done
