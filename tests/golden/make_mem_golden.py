#!/usr/bin/env python
"""Generates tests/golden/mems/*.npz: MEM mode of the UNMODIFIED reference -- SeedFinder::seeds_on_paths(sequence, cb) ->
find_mems (reference seed_finder.hpp:1459-1479, index_iter.hpp:854-906) -- on the paths the reference picked in the
same run (oracle/_ref/psi_ref_driver --paths --mems).

    make -C oracle ref && python tests/golden/make_mem_golden.py

Per case: the picked paths (ranks, head, tail), the parameters (k = minimum length, gocc threshold -r, max_mem -E), the
reads used (first max_reads of a committed reads file) and the hits as the canonical set: unique rows
(read ordinal, read offset, match length, gocc, coordinate node id, node offset).  The reference picks its paths at
random (std::random_device): every run of this script writes a different, equally valid draw."""
from __future__ import annotations

import importlib.util
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

from oracle import oracle_py as orc  # noqa: E402

spec = importlib.util.spec_from_file_location("mlg", HERE / "make_loci_golden.py")
mlg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mlg)

REF_DATA = Path("/root/reference/test/data")
X_READS, M_READS = "inputs/reads_n10000l100e0i0.fa.gz", "inputs/m_reads_n2000l100.fa.gz"
CASES = [  # name, graph, reads, k, n_paths, patched, gocc, max_mem, max_reads
    ("x_k20_n4", REF_DATA / "small/x.gfa", X_READS, 20, 4, True, 0, 0, 400),
    ("x_k12_n8_full_r5", REF_DATA / "small/x.gfa", X_READS, 12, 8, False, 5, 0, 400),      # the threshold extends matches
    ("x_k30_n2_full_r1", REF_DATA / "small/x.gfa", X_READS, 30, 2, False, 1, 0, 400),      # ... beyond 32 characters
    ("x_k16_n4_E3", REF_DATA / "small/x.gfa", X_READS, 16, 4, True, 0, 3, 400),
    ("m_k20_n4", REF_DATA / "middle/m.gfa", M_READS, 20, 4, True, 0, 0, 300),
    ("m_k24_n16_r2", REF_DATA / "middle/m.gfa", M_READS, 24, 16, True, 2, 0, 300),
    ("multi_k32_n4", REF_DATA / "multi/multi.gfa", X_READS, 32, 4, True, 0, 0, 400),
    ("fuzz02_k12_n8", HERE / "fuzz/case_02.gfa", "fuzz/case_02.fa", 12, 8, True, 0, 0, 300),    # N in graph and reads
]


def main():
    assert orc.have_reference(), "build the reference first: make -C oracle ref"
    out_dir = HERE / "mems"
    out_dir.mkdir(exist_ok=True)
    for name, gfa, reads, k, n, patched, gocc, max_mem, max_reads in CASES:
        with tempfile.TemporaryDirectory() as td:
            pf, mf, nf = (os.path.join(td, x) for x in ("paths", "mems", "nodes"))
            cmd = [os.fspath(orc.REF_DRIVER), "--gfa", os.fspath(gfa), "--fastq", os.fspath(HERE / reads), "-k", str(k), "-n", str(n),
                   "-r", str(gocc), "-E", str(max_mem), "--max-reads", str(max_reads), "--paths", pf, "--mems", mf, "--nodes", nf]
            if not patched:
                cmd.append("-P")
            subprocess.run(cmd, check=True, capture_output=True, env=dict(os.environ, OMP_NUM_THREADS="1"))
            path_ptr, ranks, head, tail, _ = mlg.parse_paths(open(pf, "rb").read())
            raw = np.fromfile(mf, "<u8").reshape(-1, 6)
        rel = os.path.relpath(gfa, HERE) if str(gfa).startswith(str(HERE)) else f"inputs/{Path(gfa).stem}.gfa.gz"
        mems = np.unique(raw, axis=0)
        np.savez_compressed(out_dir / f"{name}.npz", gfa=rel, reads=reads, k=k, n_paths=n, patched=patched, gocc=gocc, max_mem=max_mem,
                            max_reads=max_reads, path_ptr=path_ptr, nodes=ranks.astype(np.uint32), head=head, tail=tail, mems=mems,
                            n_raw=len(raw))
        print(name, "paths", len(head), "raw hits", len(raw), "unique", len(mems), "max len", int(mems[:, 2].max()) if len(mems) else 0)


if __name__ == "__main__":
    main()
