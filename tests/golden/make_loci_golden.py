#!/usr/bin/env python
"""Generates tests/golden/loci/*.npz: the paths the UNMODIFIED reference picked and the starting loci it derived from
them (SeedFinder::pick_paths + add_uncovered_loci, reference seed_finder.hpp:1138-1167,1481-1541), dumped by
oracle/_ref/psi_ref_driver --paths --loci on the reference's own fixture graphs.

    make -C oracle ref && python tests/golden/make_loci_golden.py

The reference picks paths with std::random_device, so every run of this script produces different (equally valid)
fixtures; what is committed is one such draw per case.  Per case: path_ptr / nodes (0-based ranks) / head / tail of the
picked paths (patched mode: short patches with trimmed ends), the forward text of every path as the reference's
sequence(path, Forward) returned it (checked here against the text rebuilt from ranks and trims), and the reference's
starting loci as (rank, offset).  /root/reference does not exist on the GPU box; tests only read what this wrote.
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, os.fspath(ROOT))
sys.path.insert(0, os.fspath(ROOT / "tests"))

import util  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402
from psi_b200 import capi  # noqa: E402

REF_DATA = Path("/root/reference/test/data")

X_READS = REF_DATA / "small/reads_n10000l100e0i0.fastq"
M_READS = HERE / "inputs/m_reads_n2000l100.fa.gz"
CASES = [  # name, graph (reference fixture or committed fuzz graph), k, n_paths, patched, step, gocc threshold, reads
    ("tiny_k12_n4", REF_DATA / "tiny/tiny.gfa", 12, 4, True, 1, 0, None),      # the case of test/src/test_seedfinder.cpp:85-165
    ("tiny_k12_n8", REF_DATA / "tiny/tiny.gfa", 12, 8, True, 1, 0, None),
    ("x_k12_n4", REF_DATA / "small/x.gfa", 12, 4, True, 1, 0, None),
    ("x_k20_n8", REF_DATA / "small/x.gfa", 20, 8, True, 1, 0, None),
    ("x_k20_n2_full", REF_DATA / "small/x.gfa", 20, 2, False, 1, 0, None),
    ("multi_k32_n4", REF_DATA / "multi/multi.gfa", 32, 4, True, 1, 0, None),
    ("m_k20_n4", REF_DATA / "middle/m.gfa", 20, 4, True, 1, 0, None),
    ("m_k32_n16", REF_DATA / "middle/m.gfa", 32, 16, True, 1, 0, None),
    ("fuzz02_k12_n8", HERE / "fuzz/case_02.gfa", 12, 8, True, 1, 0, None),      # graph with N bases
    ("fuzz07_k24_n4", HERE / "fuzz/case_07.gfa", 24, 4, True, 1, 0, None),
    # -e (step size) and -r (seed genome occurrence count threshold): the seed set depends on the picked paths, so the
    # same run also finds the seeds of a read sample -- the golden set for exactly these paths and loci
    ("x_k12_n4_e3", REF_DATA / "small/x.gfa", 12, 4, True, 3, 0, X_READS),
    ("m_k20_n4_e2", REF_DATA / "middle/m.gfa", 20, 4, True, 2, 0, M_READS),
    ("x_k12_n4_r1", REF_DATA / "small/x.gfa", 12, 4, True, 1, 1, X_READS),
    ("x_k12_n8_full_r5", REF_DATA / "small/x.gfa", 12, 8, False, 1, 5, X_READS),
    ("x_k20_n8_r2", REF_DATA / "small/x.gfa", 20, 8, True, 1, 2, X_READS),
    ("m_k20_n4_r2", REF_DATA / "middle/m.gfa", 20, 4, True, 1, 2, M_READS),
    ("m_k20_n16_e2_r3", REF_DATA / "middle/m.gfa", 20, 16, True, 2, 3, M_READS),
]
MAX_READS = 1500


def parse_paths(raw: bytes):
    a = np.frombuffer(raw, np.uint8)
    pos = 0

    def u64(n=1):
        nonlocal pos
        v = a[pos:pos + 8 * n].view("<u8").copy()
        pos += 8 * n
        return v
    n_paths = int(u64()[0])
    path_ptr, nodes, head, tail, texts = [0], [], [], [], []
    for _ in range(n_paths):
        n_nodes, h, t, tl = (int(x) for x in u64(4))
        nodes.append(u64(n_nodes))
        head.append(h)
        tail.append(t)
        padded = (tl + 7) // 8 * 8
        texts.append(a[pos:pos + tl].tobytes())
        pos += padded
        path_ptr.append(path_ptr[-1] + n_nodes)
    assert pos == len(a)
    return np.array(path_ptr, np.uint64), (np.concatenate(nodes) if nodes else np.zeros(0, np.uint64)), \
        np.array(head, np.uint32), np.array(tail, np.uint32), texts


def main():
    assert orc.have_reference(), "build the reference first: make -C oracle ref"
    out_dir = HERE / "loci"
    out_dir.mkdir(exist_ok=True)
    for name, gfa, k, n, patched, step, gocc, reads in CASES:
        seeds = np.zeros((0, 4), np.uint64)
        with tempfile.TemporaryDirectory() as td:
            pf, lf, nf, of = (os.path.join(td, x) for x in ("paths", "loci", "nodes", "seeds"))
            cmd = [os.fspath(orc.REF_DRIVER), "--gfa", os.fspath(gfa), "-k", str(k), "-n", str(n), "-e", str(step), "-r", str(gocc),
                   "--paths", pf, "--loci", lf, "--nodes", nf] + ([] if patched else ["-P"])
            if reads is not None:
                cmd += ["--fastq", os.fspath(reads), "--max-reads", str(MAX_READS), "-d", str(k), "--out", of]
            for attempt in range(5):     # the reference's traverser has a use-after-free that now and then kills a run
                r = subprocess.run(cmd, capture_output=True, env=dict(os.environ, OMP_NUM_THREADS="1"))
                if r.returncode == 0:
                    break
            assert r.returncode == 0, r.stderr[-500:]
            path_ptr, ids, head, tail, texts = parse_paths(open(pf, "rb").read())
            node_tab = np.fromfile(nf, "<u8").reshape(-1, 3)           # per rank: internal id, coordinate id, label length
            loci = np.fromfile(lf, "<u8").reshape(-1, 2)
            if reads is not None:
                seeds = np.fromfile(of, "<u8").reshape(-1, 4)           # canonical (read_id, read_off, coordinate node id, node_off)
        rank_of = {int(i): r for r, i in enumerate(node_tab[:, 0])}
        nodes = ids.astype(np.uint32)          # the driver dumps ranks (flattened by include/psi_b200_gum.hpp)
        loci_rank = np.array([rank_of[int(i)] for i in loci[:, 0]], np.uint32)
        loci_off = loci[:, 1].astype(np.uint32)
        # the text rebuilt from (ranks, head, tail) must be the text the reference indexed
        g = capi.Graph.load_gfa(gfa)
        assert np.array_equal(g.internal_id, node_tab[:, 0])
        for p in range(len(head)):
            seq = b"".join(g.seq[int(g.seq_start[v]):int(g.seq_start[v + 1])].tobytes() for v in nodes[int(path_ptr[p]):int(path_ptr[p + 1])])
            seq = seq[int(head[p]):len(seq) - int(tail[p])]
            assert seq == texts[p], (name, p, len(seq), len(texts[p]))
        rel = os.path.relpath(gfa, HERE) if str(gfa).startswith(str(HERE)) else f"inputs/{Path(gfa).stem}.gfa.gz"
        reads_rel = "" if reads is None else ("inputs/reads_n10000l100e0i0.fa.gz" if reads == X_READS else "inputs/m_reads_n2000l100.fa.gz")
        np.savez_compressed(out_dir / f"{name}.npz", gfa=rel, k=k, n_paths=n, patched=patched, step=step, gocc=gocc, path_ptr=path_ptr,
                            nodes=nodes, head=head, tail=tail, loci_rank=loci_rank, loci_off=loci_off, reads=reads_rel,
                            max_reads=MAX_READS if reads is not None else 0, seeds=seeds)
        print(name, "paths", len(head), "path nodes", len(nodes), "reference loci", len(loci_rank), "seeds", len(seeds))


if __name__ == "__main__":
    main()
