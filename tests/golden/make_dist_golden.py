#!/usr/bin/env python
"""Generates tests/golden/dist/*.npz: paired-end distance verification of the UNMODIFIED reference --
SeedFinder::create_distance_index(dmin, dmax, PerComponent) + verify_distance(v, o, u, p) (reference
seed_finder.hpp:1193-1317 over DiVerG's range-compressed boolean matrix powers, dindex.hpp:767-914, on Kokkos' Serial
backend) -- through oracle/_ref/psi_ref_driver --dist.

    make -C oracle ref && python tests/golden/make_dist_golden.py

Per case: the graph, the window, and pseudo-random locus pairs (xorshift, fixed seed: the file is reproducible) as rows
{rank of v, offset, rank of u, offset, answer}; ranks are 0-based positions in the reference graph's rank order."""
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

from oracle import oracle_py as orc  # noqa: E402

REF_DATA = Path("/root/reference/test/data")
CASES = [  # name, graph (reference tree or committed copy), committed graph the tests load, dmin, dmax, queries
    ("tiny_8_12", REF_DATA / "tiny/tiny.gfa", "inputs/tiny.gfa.gz", 8, 12, 3000),          # the window of test_seedfinder.cpp:231-232
    ("x_10_40", REF_DATA / "small/x.gfa", "inputs/x.gfa.gz", 10, 40, 3000),
    ("x_1_1", REF_DATA / "small/x.gfa", "inputs/x.gfa.gz", 1, 1, 2000),                    # the adjacency itself
    ("x_5_5", REF_DATA / "small/x.gfa", "inputs/x.gfa.gz", 5, 5, 2000),
    ("x_100_250", REF_DATA / "small/x.gfa", "inputs/x.gfa.gz", 100, 250, 3000),
    ("multi_20_60", REF_DATA / "multi/multi.gfa", "inputs/multi.gfa.gz", 20, 60, 3000),    # several components
    ("m_50_150", REF_DATA / "middle/m.gfa", "inputs/m.gfa.gz", 50, 150, 4000),
    ("fuzz02_10_30", HERE / "fuzz/case_02.gfa", "fuzz/case_02.gfa", 10, 30, 2000),
]


def main():
    assert orc.have_reference(), "build the reference first: make -C oracle ref"
    out_dir = HERE / "dist"
    out_dir.mkdir(exist_ok=True)
    for name, gfa, committed, dmin, dmax, nq in CASES:
        with tempfile.TemporaryDirectory() as td:
            df = os.path.join(td, "dist")
            cmd = [os.fspath(orc.REF_DRIVER), "--gfa", os.fspath(gfa), "-k", "12", "-n", "1", "--dist", df, "-m", str(dmin), "-M", str(dmax),
                   "--dist-queries", str(nq)]
            subprocess.run(cmd, check=True, capture_output=True, env=dict(os.environ, OMP_NUM_THREADS="1"))
            rows = np.fromfile(df, dtype=np.uint64).reshape(-1, 5)
        assert len(rows) == nq
        np.savez_compressed(out_dir / f"{name}.npz", gfa=committed, dmin=dmin, dmax=dmax, rows=rows.astype(np.uint32))
        print(f"{name}: {nq} pairs, {int(rows[:, 4].sum())} inside the window, {int((rows[:, 0] == rows[:, 2]).sum())} within one node")


if __name__ == "__main__":
    main()
