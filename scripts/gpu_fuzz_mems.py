#!/usr/bin/env python
"""One-off MEM-mode fuzz: find_mems on the device against oracle_py.find_mems on fresh graphs, minimum lengths on both
sides of the prefix table's 12 characters, gocc thresholds and max_mem, reads with N and with errors.
Usage (GPU box): python scripts/gpu_fuzz_mems.py FIRST LAST"""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import util
from oracle import oracle_py as orc
from psi_b200 import capi

first, last = int(sys.argv[1]), int(sys.argv[2])
t0 = time.time()
for seed in range(first, last):
    rng = np.random.default_rng(4000 + seed)
    minlen = int(rng.choice([3, 5, 8, 11, 12, 13, 16, 20, 27, 32]))
    gocc = int(rng.choice([0, 0, 1, 3, 10]))
    max_mem = int(rng.choice([0, 0, 2, 5]))
    text = util.random_bubble_gfa(8000 + seed, backbone=int(rng.integers(400, 2500)), sites=int(rng.integers(10, 200)),
                                  p_snp=float(rng.choice([0.5, 0.9])), p_ins=0.2, n_frac=float(rng.choice([0.0, 0.004])))
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "f.gfa")
        open(p, "w").write(text)
        g = capi.Graph.load_gfa(p)
    n_reads = 40
    rp, bases = util.random_walk_reads(g, n_reads, int(rng.integers(max(minlen, 14), 90)), seed=seed)
    bases = bases.copy()
    flip = rng.random(len(bases)) < 0.02
    bases[flip] = np.frombuffer(b"ACGTN", np.uint8)[rng.integers(0, 5, int(flip.sum()))]
    ps = g.pick_paths(int(rng.choice([1, 2, 4])), seed=seed)
    texts, gpos = orc.path_texts(g, ps.path_ptr, ps.nodes, ps.head_off, ps.tail_trim)
    rows = []
    for r in range(n_reads):
        for st, pl, go, ti, o in orc.find_mems(texts, bases[int(rp[r]):int(rp[r + 1])].tobytes(), minlen, gocc, max_mem):
            gp = int(gpos[ti][o])
            v = int(np.searchsorted(g.seq_start, gp, side="right") - 1)
            rows.append((int(g.coord_id[v]), gp - int(g.seq_start[v]), r, st, pl, go))
    want = np.unique(np.array(rows, np.uint64).reshape(-1, 6), axis=0)
    for packed in (0, 1):
        ctx = capi.Context(minlen, 0)
        ctx.set_option("gocc_threshold", gocc)
        ctx.set_graph(g, ids="coord")
        ctx.build_mem_index(ps)
        if packed:
            ctx.submit_chunk_packed(capi.Packed.pack(rp, bases, 0), 0)
        else:
            ctx.submit_chunk(rp, bases, 0, 0)
        got = np.unique(ctx.find_mems(max_mem), axis=0)
        assert np.array_equal(got, want), (seed, minlen, gocc, max_mem, packed, len(got), len(want))
        ctx.close()
    if seed % 10 == 0:
        print(f"seed {seed} ok (minlen {minlen}, gocc {gocc}, max_mem {max_mem}, {len(want)} MEMs) {time.time() - t0:.0f} s", flush=True)
print("MEM fuzz ok")
