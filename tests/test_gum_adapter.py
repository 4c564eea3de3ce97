"""include/psi_b200_gum.hpp -- the adapter between the reference's own graph type and the C-ABI's flat arrays -- compiled
against the REAL gum headers of the reference tree and checked against libpsi_b200's own loader.  CPU only; skipped
where /root/reference is absent (the GPU box)."""
import json
import os
import subprocess
from pathlib import Path

import pytest

import util

REF = Path("/root/reference")
GUM = REF / "ext/diverg/ext/gum"
SRC = util.ROOT / "tests" / "cpp" / "gum_adapter_check.cpp"
BIN = util.ROOT / "tests" / "cpp" / "build" / "gum_adapter_check"

pytestmark = pytest.mark.skipif(not (GUM / "include/gum/graph.hpp").exists(), reason="reference tree (gum headers) not present")


def build():
    hdr = util.ROOT / "include" / "psi_b200_gum.hpp"
    if BIN.exists() and BIN.stat().st_mtime > max(SRC.stat().st_mtime, hdr.stat().st_mtime):
        return
    BIN.parent.mkdir(exist_ok=True)
    gen = util.ROOT / "oracle" / "_ref" / "gen"          # gum's config.hpp (version macros) written by oracle/Makefile
    inc = [gen, GUM / "include", GUM / "ext/sdsl-lite/include", GUM / "ext/gfakluge/src", GUM / "ext/gfakluge/src/tinyFA",
           GUM / "ext/gfakluge/src/tinyFA/pliib", GUM / "ext/parallel-hashmap"]
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-w", "-DGUM_USER_EXCLUDE_HG"] + [f"-I{p}" for p in inc] +
                   [os.fspath(SRC), "-o", os.fspath(BIN), "-L" + os.fspath(util.ROOT / "psi_b200"), "-lpsi_b200",
                    "-Wl,-rpath," + os.fspath(util.ROOT / "psi_b200"), "-lz"], check=True)


@pytest.mark.parametrize("graph", ["tiny/tiny.gfa", "small/x.gfa", "multi/multi.gfa", "middle/m.gfa"])
def test_flatten_of_the_real_gum_graph_equals_the_library_loader(graph):
    build()
    r = subprocess.run([os.fspath(BIN), os.fspath(REF / "test/data" / graph)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["mismatches"] == 0 and info["nodes"] > 0 and info["paths"] >= 1
