#!/bin/bash
# compute-sanitizer memcheck + racecheck + synccheck over the kernels added after r02s: the distance index (rows and
# enumeration, shared and global scratch), the sliced index build, MEM mode.  Usage: bash scripts/gpu_sanitizer2.sh TAG
TAG=${1:-r02t}; OUT=gpurun_out/$TAG; mkdir -p $OUT
cat > /tmp/san2.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, util
from psi_b200 import capi
# distance index on the reference's fixture pairs
z = np.load(util.GOLDEN / "dist" / "x_10_40.npz")
g = capi.Graph.load_gfa(util.GOLDEN / str(z["gfa"]))
rows = z["rows"][:1500]
for mode, cap in ((2, 256), (1, 256), (2, 64), (1, 64)):
    ctx = capi.Context(12, 0); ctx.set_graph(g, ids="coord")
    ctx.set_option("dindex_mode", mode); ctx.set_option("dindex_list_cap", cap)
    ctx.create_distance_index(int(z["dmin"]), int(z["dmax"]))
    assert np.array_equal(ctx.verify_distance(rows[:, :4]), rows[:, 4].astype(bool))
    ctx.close()
    print("distance", mode, cap, "ok", flush=True)
# sliced index build, all result formats
case = {c["name"]: c for c in util.golden_index()["cases"]}["x_k12"]
rp, bases = util.read_fasta(util.GOLDEN / case["reads"])
n = 1500; rp = rp[:n + 1]; bases = bases[:int(rp[-1])]
ref = None
for slices in (1, 16):
    ctx = capi.Context(case["k"], 0); ctx.set_option("build_slices", slices)
    ctx.set_graph(g, ids="coord"); ctx.set_paths(g.pick_paths(6, seed=2)); ctx.find_loci()
    ctx.submit_chunk(rp, bases, 0, case["d"]); a = ctx.seeds_all(); rec = capi.canonical(ctx.fetch())
    c = ctx.seeds_all(capi.ALL | capi.DENSE); d, e = ctx.fetch_dense()
    r2, _ = capi.dense_to_records(d, e, rp, case["k"], case["d"], 0)
    assert c == a and np.array_equal(capi.canonical(r2), rec)
    assert ctx.seeds_all(capi.ALL | capi.DENSE5) == a
    d5, e5 = ctx.fetch_dense5()
    assert np.array_equal(d5, d)
    assert ref is None or np.array_equal(ref, rec)
    ref = rec
    ctx.close()
    print("slices", slices, "hits", a, flush=True)
# MEM mode
zm = np.load(util.GOLDEN / "mems" / "x_k20_n4.npz")
ctx = capi.Context(int(zm["k"]), 0); ctx.set_graph(g, ids="coord")
ctx.build_mem_index(capi.PathSet(path_ptr=zm["path_ptr"], nodes=zm["nodes"], head_off=zm["head"], tail_trim=zm["tail"]))
ctx.submit_chunk(rp[:201], bases[:int(rp[200])], 0, 0)
print("mems", len(ctx.find_mems(0)), flush=True)
ctx.close()
PY
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san2.py > $OUT/${tool}_new_kernels.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/${tool}_new_kernels.log | tail -1)"
done
