#!/usr/bin/env python
"""One-off extended fuzz (not part of the suite): the suite's fuzz tests with other seeds, plus sliced builds, 5-byte
results and distance queries on every fresh graph.  Usage (GPU box): python scripts/gpu_fuzz_extra.py FIRST LAST"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import util
from oracle import oracle_py as orc
from psi_b200 import capi
import test_gpu_parity as T

first, last = int(sys.argv[1]), int(sys.argv[2])
t0 = time.time()
for seed in range(first, last):
    T.test_fuzz_fresh_graphs_all_routes_vs_oracle(seed)
    rng = np.random.default_rng(9000 + seed)
    k = int(rng.choice([8, 12, 16, 20, 24, 28, 32]))
    d = int(rng.choice([1, k // 2, k]))
    text = util.random_bubble_gfa(7000 + seed, backbone=int(rng.integers(2000, 8000)), sites=int(rng.integers(50, 600)),
                                  p_snp=float(rng.choice([0.5, 0.8, 1.0])), p_ins=0.15, n_frac=float(rng.choice([0.0, 0.003])),
                                  multi_allele=float(rng.choice([0.0, 0.3])))
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "f.gfa")
        open(p, "w").write(text)
        g = capi.Graph.load_gfa(p)
    rp, bases = util.random_walk_reads(g, 500, int(rng.integers(max(k, 40), 140)), seed=seed)
    og = orc.OGraph.of(g)
    want, _ = orc.seeds_closed_form(og, orc.OReads(rp, bases, 0), k, d)
    ps = g.pick_paths(int(rng.choice([1, 3, 6])), seed=seed)
    for slices in (1, int(rng.choice([4, 16, 64]))):
        if 2 * k < {1: 0, 4: 2, 16: 4, 64: 6}[slices]:
            continue
        ctx = capi.Context(k, 0)
        ctx.set_option("build_slices", slices)
        ctx.set_option("build_group_windows", int(rng.choice([0, 2000])))
        ctx.set_graph(g, ids="coord")
        ctx.set_paths(ps)
        ctx.find_loci()
        ctx.submit_chunk_packed(capi.Packed.pack(rp, bases, 0), d)
        n = ctx.seeds_all(capi.ALL | capi.DENSE)
        dense, extra = ctx.fetch_dense()
        rec, _ = capi.dense_to_records(dense, extra, rp, k, d, 0)
        assert n == len(rec) and np.array_equal(capi.canonical(rec), want), (seed, k, d, slices)
        assert ctx.dense5_layout()[1]
        assert ctx.seeds_all(capi.ALL | capi.DENSE5) == n
        d5, e5 = ctx.fetch_dense5()
        assert np.array_equal(d5, dense), (seed, "dense5")
        ctx.close()
    # distance queries on the same graph
    start = np.asarray(g.seq_start, np.int64)
    total = int(start[-1])
    dmin = int(rng.integers(1, 60))
    dmax = dmin + int(rng.integers(0, 150))
    a = rng.integers(0, total, 600)
    b = np.minimum(total - 1, a + rng.integers(0, 2 * dmax + 10, 600))
    rv, ru = np.searchsorted(start, a, side="right") - 1, np.searchsorted(start, b, side="right") - 1
    pairs = np.stack([rv, a - start[rv], ru, b - start[ru]], axis=1).astype(np.uint32)
    wantd = np.array([orc.verify_distance(og, int(v), int(o), int(u), int(p_), dmin, dmax) for v, o, u, p_ in pairs])
    for mode, cap in T.DIST_MODES:
        ctx = capi.Context(12, 0)
        ctx.set_graph(g, ids="coord")
        ctx.set_option("dindex_mode", mode)
        ctx.set_option("dindex_list_cap", cap)
        ctx.create_distance_index(dmin, dmax)
        assert np.array_equal(ctx.verify_distance(pairs), wantd), (seed, dmin, dmax, mode, cap)
        ctx.close()
    print(f"seed {seed} ok (k={k} d={d} window {dmin}..{dmax}) {time.time() - t0:.0f} s", flush=True)
print("extended fuzz ok")
