#!/usr/bin/env python
"""Generates tests/golden/ from the UNMODIFIED reference compiled in this container.

    make -C oracle ref            # builds oracle/_ref/psi_ref_driver from /root/reference
    python tests/golden/make_golden.py

What is written (all small, committed):
  inputs/*.gfa.gz, inputs/*.fa(.gz)   the reference's own test fixtures (test/data/{tiny,small,multi,middle})
                                      re-encoded (reads as FASTA, gzip) -- data, not source;
  fuzz/case_NN.gfa, case_NN.fa        seeded random bubble graphs + random-walk reads (tests/util.py);
  nodes_<graph>.npy                   gum's rank -> (internal id, coordinate id, label length) per fixture graph;
  golden.json                         per case: parameters, size and md5 of the canonical seed set
                                      (sorted unique (read_id, read_offset, coordinate node id, node_offset),
                                      4 x u64 LE) and the reference's raw hit counts.

/root/reference does not exist on the GPU box; tests only read what this script wrote.
"""
from __future__ import annotations

import gzip
import json
import os
import shutil
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


def gz_copy(src, dst):
    with open(src, "rb") as f, gzip.GzipFile(dst, "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)


def fastq_to_fasta(src, dst, gz):
    rp, bases = util.read_fasta(src)
    if gz:
        with gzip.GzipFile(dst, "wb", mtime=0) as g:
            for i in range(len(rp) - 1):
                g.write(b">r%d\n" % i + bases[int(rp[i]):int(rp[i + 1])].tobytes() + b"\n")
    else:
        util.write_fasta(dst, rp, bases)


def main():
    assert orc.have_reference(), "build the reference first: make -C oracle ref"
    (HERE / "inputs").mkdir(exist_ok=True)
    (HERE / "fuzz").mkdir(exist_ok=True)
    tmp = Path(tempfile.mkdtemp())
    cases = []

    graphs = {"tiny": "tiny/tiny.gfa", "x": "small/x.gfa", "multi": "multi/multi.gfa", "m": "middle/m.gfa"}
    for name, rel in graphs.items():
        gz_copy(REF_DATA / rel, HERE / "inputs" / f"{name}.gfa.gz")
        nodes = tmp / f"{name}.nodes"
        orc.run_reference(REF_DATA / rel, None, 12, n_paths=0, want_out=False, extra=["--nodes", os.fspath(nodes)])
        np.save(HERE / f"nodes_{name}.npy", np.fromfile(nodes, np.uint64).reshape(-1, 3))
    fastq_to_fasta(REF_DATA / "small/reads_n10000l100e0i0.fastq", HERE / "inputs" / "reads_n10000l100e0i0.fa.gz", True)
    fastq_to_fasta(REF_DATA / "small/reads_n10l10e0i0.fastq", HERE / "inputs" / "reads_n10l10e0i0.fa", False)

    # random-walk reads on the variant-dense `middle` graph (regenerated with tests/util.py)
    gm = capi.Graph.load_gfa(REF_DATA / "middle/m.gfa")
    rp, bases = util.random_walk_reads(gm, 2000, 100, seed=42)
    with gzip.GzipFile(HERE / "inputs" / "m_reads_n2000l100.fa.gz", "wb", mtime=0) as g:
        for i in range(len(rp) - 1):
            g.write(b">r%d\n" % i + bases[int(rp[i]):int(rp[i + 1])].tobytes() + b"\n")

    def add_case(name, gfa_src, gfa_rel, reads_src, reads_rel, k, d, n, patched=True, chunk=0, survey_md5=None):
        tuples, stats = orc.run_reference(gfa_src, reads_src, k, d=d, n_paths=n, patched=patched, chunk=chunk)
        md5 = util.md5_tuples(tuples)
        if survey_md5:
            assert md5 == survey_md5, (name, md5, survey_md5)
        cases.append({"name": name, "gfa": gfa_rel, "reads": reads_rel, "k": k, "d": d, "n_paths": n,
                      "patched": patched, "chunk": chunk, "count": int(len(tuples)), "md5": md5,
                      "ref_raw_on": stats["raw_on"], "ref_raw_off": stats["raw_off"], "ref_loci": stats["loci"],
                      "query_seeds": stats["query_seeds"]})
        print(cases[-1])

    x, xr = REF_DATA / "small/x.gfa", REF_DATA / "small/reads_n10000l100e0i0.fastq"
    # the four portable goldens of SURVEY 8c (md5s recorded there are re-checked here)
    add_case("x_k12", x, "inputs/x.gfa.gz", xr, "inputs/reads_n10000l100e0i0.fa.gz", 12, 12, 16,
             survey_md5="4baecc78144d8894ecfd5a2aade66efc")
    add_case("x_k20_c1000", x, "inputs/x.gfa.gz", xr, "inputs/reads_n10000l100e0i0.fa.gz", 20, 20, 8, chunk=1000,
             survey_md5="8bd587e67333b8e105e86625efaaabf4")
    add_case("x_k20_d1", x, "inputs/x.gfa.gz", xr, "inputs/reads_n10000l100e0i0.fa.gz", 20, 1, 2,
             survey_md5="1e51e50b0c937e338a3cd8683a0c0744")
    add_case("multi_k32", REF_DATA / "multi/multi.gfa", "inputs/multi.gfa.gz", xr,
             "inputs/reads_n10000l100e0i0.fa.gz", 32, 32, 4, chunk=2500,
             survey_md5="ddb5858e6e72b03ab74097d5e8f254bf")
    add_case("x_k12_unpatched", x, "inputs/x.gfa.gz", xr, "inputs/reads_n10000l100e0i0.fa.gz", 12, 12, 8,
             patched=False)
    add_case("x_k10_tiny_reads", x, "inputs/x.gfa.gz", REF_DATA / "small/reads_n10l10e0i0.fastq",
             "inputs/reads_n10l10e0i0.fa", 10, 10, 4)
    add_case("m_k20", REF_DATA / "middle/m.gfa", "inputs/m.gfa.gz", HERE / "inputs" / "m_reads_n2000l100.fa.gz",
             "inputs/m_reads_n2000l100.fa.gz", 20, 20, 4)
    add_case("m_k32", REF_DATA / "middle/m.gfa", "inputs/m.gfa.gz", HERE / "inputs" / "m_reads_n2000l100.fa.gz",
             "inputs/m_reads_n2000l100.fa.gz", 32, 32, 16)

    # fuzz: random bubble graphs, 300 random-walk 70 bp reads
    fuzz_params = []
    i = 0
    for k in (12, 16, 24, 32):
        for n, d in ((2, 0), (4, 1), (8, 0)):
            fuzz_params.append((i, k, n, d if d else k, 2500 + 500 * (i % 5), 60 + 35 * (i % 7), i % 3 == 2))
            i += 1
    for (i, k, n, d, backbone, sites, with_n) in fuzz_params:
        gfa = HERE / "fuzz" / f"case_{i:02d}.gfa"
        fa = HERE / "fuzz" / f"case_{i:02d}.fa"
        gfa.write_text(util.random_bubble_gfa(1000 + i, backbone=backbone, sites=sites,
                                              n_frac=0.002 if with_n else 0.0))
        g = capi.Graph.load_gfa(gfa)
        rp, bases = util.random_walk_reads(g, 300, 70, seed=2000 + i, n_frac=0.003 if with_n else 0.0)
        util.write_fasta(fa, rp, bases)
        add_case(f"fuzz_{i:02d}", gfa, f"fuzz/{gfa.name}", fa, f"fuzz/{fa.name}", k, d, n, chunk=150)

    with open(HERE / "golden.json", "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "reference": "cartoonist/psi @ /root/reference "
                   "(compiled in place by oracle/Makefile)", "cases": cases}, f, indent=1)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
