#!/usr/bin/env python
"""Generates tests/golden/refindex/: path indexes SAVED BY THE UNMODIFIED REFERENCE (SeedFinder::serialize_path_index,
what `psikt -I <prefix>` writes; oracle/_ref/psi_ref_driver --save-index) -- the files `<name>_paths` and
`<name>_loci_e1l<k>` byte for byte as the reference wrote them -- plus, per case, `<name>.npz` with the same paths as
flat arrays (the driver's --paths dump of the same run) for the loader test.  The reference's `<name>` file itself (its
serialised csa_wt) is not kept: this build rebuilds its device index from the paths.

    make -C oracle ref && python tests/golden/make_refindex_golden.py
"""
from __future__ import annotations

import importlib.util
import os
import shutil
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
CASES = [  # name, graph, k, n_paths, patched
    ("x_k12_n16", REF_DATA / "small/x.gfa", 12, 16, True),       # the graph / k / n of the golden case x_k12
    ("x_k20_n8", REF_DATA / "small/x.gfa", 20, 8, True),
    ("multi_k32_n4", REF_DATA / "multi/multi.gfa", 32, 4, True),
    ("m_k20_n4", REF_DATA / "middle/m.gfa", 20, 4, True),        # paths of ~3 400 nodes: many enc_vector samples
    ("fuzz02_k12_n3_full", HERE / "fuzz/case_02.gfa", 12, 3, False),
]


def main():
    assert orc.have_reference(), "build the reference first: make -C oracle ref"
    out_dir = HERE / "refindex"
    out_dir.mkdir(exist_ok=True)
    for name, gfa, k, n, patched in CASES:
        with tempfile.TemporaryDirectory() as td:
            prefix, pf = os.path.join(td, name), os.path.join(td, "paths.bin")
            cmd = [os.fspath(orc.REF_DRIVER), "--gfa", os.fspath(gfa), "-k", str(k), "-n", str(n), "--save-index", prefix, "--paths", pf]
            if not patched:
                cmd.append("-P")
            subprocess.run(cmd, check=True, capture_output=True, env=dict(os.environ, OMP_NUM_THREADS="1"))
            path_ptr, ranks, head, tail, _ = mlg.parse_paths(open(pf, "rb").read())
            shutil.copy(prefix + "_paths", out_dir / f"{name}_paths")
            shutil.copy(prefix + f"_loci_e1l{k}", out_dir / f"{name}_loci_e1l{k}")
            loci = np.fromfile(prefix + f"_loci_e1l{k}", "<u8")
        rel = os.path.relpath(gfa, HERE) if str(gfa).startswith(str(HERE)) else f"inputs/{Path(gfa).stem}.gfa.gz"
        np.savez_compressed(out_dir / f"{name}.npz", gfa=rel, k=k, n_paths=n, patched=patched, path_ptr=path_ptr,
                            nodes=ranks.astype(np.uint32), head=head, tail=tail, n_loci=int(loci[0]))
        print(name, "paths", len(head), "nodes", len(ranks), "loci", int(loci[0]), "bytes", os.path.getsize(out_dir / f"{name}_paths"))


if __name__ == "__main__":
    main()
