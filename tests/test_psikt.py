"""The reference-facing C++ layer: psi::SeedFinder (psi_b200/include/psi/seed_finder.hpp) and the
psikt CLI (psi_b200/src/psikt.cpp).  CPU tests cover option parsing, help text and the loud
failure without a device; GPU tests drive the CLI and the API the way the reference's
find_seeds() does (reference src/psikt.cpp:83-212) and compare bit-exactly with the goldens."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import util
from psi_b200 import capi

ROOT = util.ROOT
PSIKT = ROOT / "psi_b200" / "bin" / "psikt"
DRIVER_SRC = ROOT / "tests" / "cpp" / "seed_finder_driver.cpp"
DRIVER = ROOT / "tests" / "cpp" / "build" / "seed_finder_driver"
G = {c["name"]: c for c in util.golden_index()["cases"]}


def run(cmd, **kw):
    return subprocess.run([str(c) for c in cmd], capture_output=True, text=True, timeout=600, **kw)


def build_driver(src=DRIVER_SRC, exe=DRIVER):
    hdrs = list((ROOT / "psi_b200" / "include" / "psi").glob("*.hpp"))
    if exe.exists() and exe.stat().st_mtime > max(f.stat().st_mtime for f in [src] + hdrs):
        return
    exe.parent.mkdir(exist_ok=True)
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-Wall", "-o", os.fspath(exe), os.fspath(src),
                    "-L" + os.fspath(ROOT / "psi_b200"), "-lpsi_b200", "-lpthread", "-Wl,-rpath," + os.fspath(ROOT / "psi_b200")], check=True)


DIST_DRIVER_SRC = ROOT / "tests" / "cpp" / "distance_driver.cpp"
DIST_DRIVER = ROOT / "tests" / "cpp" / "build" / "distance_driver"


def gunzip_to(src, dst):
    import gzip
    with gzip.open(src, "rb") as f, open(dst, "wb") as o:
        o.write(f.read())
    return dst


# ------------------------------------------------------------------ CPU --

def test_psikt_exists_and_help_lists_the_reference_option_table():
    assert PSIKT.exists(), "psi_b200/bin/psikt missing: run __graft_entry__.build()"
    r = run([PSIKT, "--help"])
    assert r.returncode == 0
    # reference src/psikt.cpp:306-435
    for short, long_ in [("f", "fastq"), ("o", "output"), ("I", "path-index"), ("l", "seed-length"), ("c", "chunk-size"),
                         ("e", "step-size"), ("d", "distance"), ("n", "path-num"), ("P", "no-patched"), ("t", "context"),
                         ("r", "gocc-threshold"), ("E", "max-mem"), ("m", "min-insert-size"), ("M", "max-insert-size"),
                         ("i", "index"), ("x", "index-only"), ("L", "log-file"), ("Q", "no-log-file"), ("q", "quiet"),
                         ("C", "no-color"), ("D", "disable-log"), ("v", "verbose")]:
        assert f"-{short}, --{long_}" in r.stdout, long_
    assert "--dindex-mode" in r.stdout


@pytest.mark.parametrize("args,msg", [
    (["-l", "12", "g.gfa"], "--fastq is required"),
    (["-f", "r.fq", "g.gfa"], "--seed-length is required"),
    (["-f", "r.fq", "-l", "12"], "Not enough arguments"),
    (["-f", "r.fq", "-l", "12", "a.gfa", "b.gfa"], "Too many arguments"),
    (["-f", "r.fq", "-l", "twelve", "g.gfa"], "cannot be casted to integer"),
    (["-f", "r.fq", "-l", "12", "g.txt"], "invalid file extension"),
    (["-f", "r.txt", "-l", "12", "g.gfa"], "invalid file extension"),
    (["-f", "r.fq", "-l", "12", "-i", "BWT", "g.gfa"], "Undefined index type"),
    (["-f", "r.fq", "-l", "12", "--dindex-mode", "fast", "g.gfa"], "allowed values"),
    (["-f", "r.fq", "-l", "12", "--bogus", "g.gfa"], "unknown option"),
])
def test_psikt_argument_errors_exit_1(args, msg):
    r = run([PSIKT] + args)
    assert r.returncode == 1
    assert msg in r.stderr


def test_psikt_missing_inputs_fail_with_the_reference_messages(tmp_path):
    gfa = gunzip_to(util.GOLDEN / "inputs/tiny.gfa.gz", tmp_path / "tiny.gfa")
    r = run([PSIKT, "-f", tmp_path / "nope.fq", "-l", "10", "-Q", "-q", "-o", tmp_path / "o", gfa])
    assert r.returncode == 1 and "could not open file" in r.stderr
    r = run([PSIKT, "-f", util.GOLDEN / "inputs/reads_n10l10e0i0.fa", "-l", "10", "-Q", "-q", "-o", tmp_path / "o", tmp_path / "missing.gfa"])
    assert r.returncode == 1 and "could not open" in r.stderr


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="checks the failure mode of a box without a GPU")
def test_psikt_fails_loudly_without_a_gpu(tmp_path):
    gfa = gunzip_to(util.GOLDEN / "inputs/tiny.gfa.gz", tmp_path / "tiny.gfa")
    r = run([PSIKT, "-f", util.GOLDEN / "inputs/reads_n10l10e0i0.fa", "-l", "10", "-n", "2", "-Q", "-q", "-o", tmp_path / "o", gfa])
    assert r.returncode == 1
    assert "no CPU fallback" in r.stderr


def test_mirror_headers_compile_standalone(tmp_path):
    """The SeedFinder mirror is host C++17 over the C-ABI only: it must compile with g++ alone."""
    src = tmp_path / "t.cpp"
    src.write_text('#include "%s"\nint main() { return sizeof(psi::SeedFinder<>) > 0 ? 0 : 1; }\n'
                   % os.fspath(ROOT / "psi_b200/include/psi/seed_finder.hpp"))
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", os.fspath(src)], check=True)
    # the callers written against the reference's API (find_seeds of the CLI, the reference's distance-index scenario,
    # the lifecycle knobs) instantiate every method they use: they must compile on a box without a GPU too
    for caller in ("tests/cpp/seed_finder_driver.cpp", "tests/cpp/distance_driver.cpp", "psi_b200/src/psikt.cpp"):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", os.fspath(ROOT / caller)], check=True)


# ------------------------------------------------------------------ GPU --

def load_psikt_output(path, g):
    rec = np.fromfile(path, dtype="<u8").reshape(-1, 4)
    # psikt writes gum's internal ids: map back to the coordinate ids of the goldens
    order = np.argsort(g.internal_id)
    idx = order[np.searchsorted(g.internal_id[order], rec[:, 0])]
    assert np.array_equal(g.internal_id[idx], rec[:, 0])
    rec = rec.copy()
    rec[:, 0] = g.coord_id[idx]
    return rec


@pytest.mark.gpu
@pytest.mark.parametrize("name,extra", [("x_k12", []), ("x_k20_c1000", ["-c", "1000"]), ("multi_k32", ["-c", "2500"]),
                                        ("x_k20_d1", []), ("m_k20", ["-P"])])
def test_psikt_output_is_the_reference_seed_set(tmp_path, name, extra):
    c = G[name]
    gfa = gunzip_to(util.GOLDEN / c["gfa"], tmp_path / "g.gfa") if c["gfa"].endswith(".gz") else util.GOLDEN / c["gfa"]
    reads = util.GOLDEN / c["reads"]
    if not any(str(reads).endswith(e) for e in (".fa", ".fa.gz", ".fq", ".fq.gz", ".fasta", ".fastq")):
        pytest.skip("reads extension")
    out, log = tmp_path / "seeds.bin", tmp_path / "psi.log"
    r = run([PSIKT, "-f", reads, "-l", c["k"], "-d", c["d"], "-n", c["n_paths"], "-o", out, "-L", log, "-q"] + extra + [gfa])
    assert r.returncode == 0, r.stderr
    g = capi.Graph.load_gfa(gfa)
    rec = load_psikt_output(out, g)
    got = capi.canonical(rec)
    assert len(rec) == len(got) == c["count"], "psikt writes each hit of the set exactly once"
    assert util.md5_tuples(got) == c["md5"]
    text = log.read_text()
    for line in ["Parameters:", "- Seed length: %d" % c["k"], "Loading input graph from file", "Number of starting loci (in ",
                 "Finding seeds...", "Fetched ", "Seeding done in ", "Found seeds on paths in ", "Found seeds off paths in ",
                 "Total number of seeds found: %d" % c["count"], "-> of which found off paths: ", "Total number of reads covered: ",
                 "Total number of 'godown' operations: ", "All Timers"]:
        assert line in text, line
    covered = len(np.unique(got[:, 0]))
    assert "Total number of reads covered: %d" % covered in text


@pytest.mark.gpu
def test_psikt_saved_path_index_is_reloaded(tmp_path):
    c = G["x_k12"]
    gfa = gunzip_to(util.GOLDEN / c["gfa"], tmp_path / "g.gfa")
    reads = util.GOLDEN / c["reads"]
    prefix = tmp_path / "idx"
    r = run([PSIKT, "-f", reads, "-l", c["k"], "-n", "4", "-I", prefix, "-x", "-L", tmp_path / "a.log", "-q", "-o", tmp_path / "o0", gfa])
    assert r.returncode == 0, r.stderr
    assert "Skipping seed finding as requested" in (tmp_path / "a.log").read_text()
    loci_file = Path(str(prefix) + "_loci_e1l%d" % c["k"])
    assert Path(str(prefix) + "_paths.b200").exists() and loci_file.exists()
    # the loci file has the reference's byte layout: u64 count, then {i64 coordinate id, u64 offset} (seed_finder.hpp:1659-1679)
    raw = np.fromfile(loci_file, dtype="<u8")
    assert raw[0] == (len(raw) - 1) // 2
    r = run([PSIKT, "-f", reads, "-l", c["k"], "-I", prefix, "-L", tmp_path / "b.log", "-q", "-o", tmp_path / "o1", gfa])
    assert r.returncode == 0, r.stderr
    text = (tmp_path / "b.log").read_text()
    assert "The path index has been found and loaded." in text
    g = capi.Graph.load_gfa(gfa)
    got = capi.canonical(load_psikt_output(tmp_path / "o1", g))
    assert util.md5_tuples(got) == c["md5"]
    # a damaged paths file, or one written for another graph, is refused and the index is rebuilt
    pf = Path(str(prefix) + "_paths.b200")
    good = pf.read_bytes()
    for bad in (good[:-7], good[:40] + b"\xff" * 8 + good[48:], good[:64] + b"\xff\xff\xff\x7f" + good[68:]):
        pf.write_bytes(bad)
        r = run([PSIKT, "-f", reads, "-l", c["k"], "-n", "4", "-I", prefix, "-L", tmp_path / "d.log", "-q", "-o", tmp_path / "o3", gfa])
        assert r.returncode == 0, r.stderr
        assert "No valid path index found. Creating the path index..." in (tmp_path / "d.log").read_text()
        assert util.md5_tuples(capi.canonical(load_psikt_output(tmp_path / "o3", g))) == c["md5"]
        assert pf.read_bytes() != bad                  # rewritten by the rebuild
    # insert sizes (-m / -M): the distance index is created with the path index, noted beside it, and found again
    p2 = tmp_path / "idx2"
    r = run([PSIKT, "-f", reads, "-l", c["k"], "-n", "4", "-I", p2, "-m", "100", "-M", "300", "-x", "-L", tmp_path / "e.log", "-q", "-o",
             tmp_path / "o4", gfa])
    assert r.returncode == 0, r.stderr
    text = (tmp_path / "e.log").read_text()
    assert "Constructing distance index" in text and "Created distance index in" in text and "Saved distance index in" in text
    assert Path(str(p2) + "_dist_mat_m100M300.b200").exists()
    r = run([PSIKT, "-f", reads, "-l", c["k"], "-I", p2, "-m", "100", "-M", "300", "-L", tmp_path / "f.log", "-q", "-o", tmp_path / "o5", gfa])
    assert r.returncode == 0, r.stderr
    assert "The path index has been found and loaded." in (tmp_path / "f.log").read_text()
    assert util.md5_tuples(capi.canonical(load_psikt_output(tmp_path / "o5", g))) == c["md5"]
    # -n 0 and no index: nothing is found (src/psikt.cpp:121-123)
    r = run([PSIKT, "-f", reads, "-l", c["k"], "-L", tmp_path / "c.log", "-q", "-o", tmp_path / "o2", gfa])
    assert r.returncode == 0
    assert os.path.getsize(tmp_path / "o2") == 0
    assert "No path has been specified. Skipping path indexing..." in (tmp_path / "c.log").read_text()


@pytest.mark.gpu
def test_psikt_loads_a_path_index_saved_by_the_reference(tmp_path):
    """`psikt -I <prefix>` with the files the REFERENCE's psikt -I wrote (tests/golden/refindex: `<prefix>_paths` and
    `<prefix>_loci_e1l<k>`, byte for byte): the index is found and loaded -- paths decoded from sdsl's enc_vector, loci in
    the shared format -- and the seed set is the golden one."""
    import shutil
    c = G["x_k12"]
    gfa = gunzip_to(util.GOLDEN / c["gfa"], tmp_path / "g.gfa")
    for f in ("x_k12_n16_paths", "x_k12_n16_loci_e1l12"):
        shutil.copy(util.GOLDEN / "refindex" / f, tmp_path / f.replace("x_k12_n16", "idx"))
    r = run([PSIKT, "-f", util.GOLDEN / c["reads"], "-l", c["k"], "-I", tmp_path / "idx", "-L", tmp_path / "a.log", "-q", "-o", tmp_path / "o", gfa])
    assert r.returncode == 0, r.stderr
    text = (tmp_path / "a.log").read_text()
    assert "The path index has been found and loaded." in text and "No valid path index found" not in text
    z = np.load(util.GOLDEN / "refindex" / "x_k12_n16.npz")
    assert "Total number of starting loci: %d" % int(z["n_loci"]) in text      # the reference's own loci were loaded
    g = capi.Graph.load_gfa(gfa)
    assert util.md5_tuples(capi.canonical(load_psikt_output(tmp_path / "o", g))) == c["md5"]
    # a shorter seed length: the paths are loaded, the loci (none saved for this k) are computed and saved beside them
    c10 = G["x_k10_tiny_reads"]
    r = run([PSIKT, "-f", util.GOLDEN / c10["reads"], "-l", "10", "-I", tmp_path / "idx", "-L", tmp_path / "b.log", "-q", "-o", tmp_path / "o2", gfa])
    assert r.returncode == 0, r.stderr
    assert "The path index has been found and loaded." in (tmp_path / "b.log").read_text() and (tmp_path / "idx_loci_e1l10").exists()
    assert util.md5_tuples(capi.canonical(load_psikt_output(tmp_path / "o2", g))) == c10["md5"]
    # a seed longer than the context the index was patched with is refused, as by the reference (seed_finder.hpp:1434-1437)
    r = run([PSIKT, "-f", util.GOLDEN / c["reads"], "-l", "20", "-I", tmp_path / "idx", "-Q", "-q", "-o", tmp_path / "o3", gfa])
    assert r.returncode == 1 and "seed length should not be larger than context size" in r.stderr


@pytest.mark.gpu
def test_psikt_graph_without_embedded_path_is_an_error(tmp_path):
    gfa = tmp_path / "nopath.gfa"
    gfa.write_text("H\tVN:Z:1.0\nS\t1\tACGTACGTACGTAAACCCGGGTTT\nS\t2\tACGT\nL\t1\t+\t2\t+\t0M\n")
    r = run([PSIKT, "-f", util.GOLDEN / "inputs/reads_n10l10e0i0.fa", "-l", "10", "-n", "2", "-Q", "-q", "-o", tmp_path / "o", gfa])
    assert r.returncode == 1
    assert "no reference path found in the input graph" in r.stderr      # seed_finder.hpp:1145-1147


@pytest.mark.gpu
@pytest.mark.parametrize("name,chunk", [("x_k12", 0), ("m_k20", 700), ("fuzz_03", 150)])
def test_seed_finder_api_matches_oracle(tmp_path, name, chunk):
    from oracle import oracle_py as orc
    build_driver()
    c = G[name]
    gfa = gunzip_to(util.GOLDEN / c["gfa"], tmp_path / "g.gfa") if c["gfa"].endswith(".gz") else util.GOLDEN / c["gfa"]
    r = run([DRIVER, gfa, util.GOLDEN / c["reads"], c["k"], c["d"], c["n_paths"], chunk, tmp_path / "out"])
    assert r.returncode == 0, r.stdout + r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["errors_ok"] and info["infos"] >= 4

    def load(ext):
        return np.fromfile(str(tmp_path / "out") + ext, dtype="<u8").reshape(-1, 4)
    on, off, all1, all2, mt = load(".on"), load(".off"), load(".all1"), load(".all2"), load(".mt")
    assert len(mt) == c["count"] and util.md5_tuples(np.unique(mt, axis=0)) == c["md5"], "three threads over one const finder"
    for part in (on, off, all1, all2):
        assert len(np.unique(part, axis=0)) == len(part), "callbacks see each hit once"
    sep = np.unique(np.concatenate([on, off]), axis=0)
    both = np.unique(np.concatenate([all1, all2]), axis=0)
    assert len(sep) == len(on) + len(off) == c["count"]
    assert util.md5_tuples(sep) == c["md5"] == util.md5_tuples(both)
    # seeds_all hands the on-path hits to callback1 and the rest to callback2 (seed_finder.hpp:1734-1743)
    assert np.array_equal(np.unique(on, axis=0), np.unique(all1, axis=0))
    assert np.array_equal(np.unique(off, axis=0), np.unique(all2, axis=0))
    g = capi.Graph.load_gfa(gfa)
    # MEM mode through the mirror (seeds_on_paths(sequence, cb)): the restatement of find_mems on the very paths the
    # mirror picked (psi_b200_pick_paths is seeded, so they can be drawn again here)
    mems = np.fromfile(str(tmp_path / "out") + ".mems", dtype="<u8").reshape(-1, 6)
    ps = g.pick_paths(c["n_paths"], True, c["k"], seed=0x9e3779b97f4a7c15)
    texts, gpos = orc.path_texts(g, ps.path_ptr, ps.nodes, ps.head_off, ps.tail_trim)
    rp, bases = util.read_fasta(util.GOLDEN / c["reads"])
    rows = []
    for r in range(min(40, len(rp) - 1, chunk or 40)):
        for st, pl, go, ti, o in orc.find_mems(texts, bases[int(rp[r]):int(rp[r + 1])].tobytes(), c["k"]):
            gp = int(gpos[ti][o])
            v = int(np.searchsorted(g.seq_start, gp, side="right") - 1)
            rows.append((r, st, pl, go, int(g.coord_id[v]), gp - int(g.seq_start[v])))
    want = np.unique(np.array(rows, np.uint64).reshape(-1, 6), axis=0)
    assert len(mems) == len(want) and np.array_equal(np.unique(mems, axis=0), want)
    loci = np.fromfile(str(tmp_path / "out") + ".loci", dtype="<u8").reshape(-1, 2)
    assert info["loci"] == len(loci) and info["uniq_nodes"] == len(np.unique(loci[:, 0]))
    assert set(loci[:, 0].tolist()) <= set(g.coord_id.tolist())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny_8_12", "x_10_40", "multi_20_60", "m_50_150"])
def test_seed_finder_distance_api_matches_the_reference(tmp_path, name):
    """The reference's own scenario (test/src/test_seedfinder.cpp:225-312) through the mirror: create_distance_index,
    verify_distance one by one, in bulk and from three threads, save_distance_index, a second finder that opens the saved
    index -- every answer equal to the one the compiled reference gave (tests/golden/dist)."""
    build_driver(DIST_DRIVER_SRC, DIST_DRIVER)
    z = np.load(util.GOLDEN / "dist" / f"{name}.npz")
    gz = util.GOLDEN / str(z["gfa"])
    gfa = gunzip_to(gz, tmp_path / "g.gfa") if str(gz).endswith(".gz") else gz
    g = capi.Graph.load_gfa(gfa)
    rows = z["rows"].astype(np.uint64)
    pairs = rows[:, :4].copy()
    pairs[:, 0] = g.coord_id[rows[:, 0]]
    pairs[:, 2] = g.coord_id[rows[:, 2]]
    pairs.astype("<u8").tofile(tmp_path / "pairs")
    r = run([DIST_DRIVER, gfa, int(z["dmin"]), int(z["dmax"]), tmp_path / "pairs", tmp_path / "out", tmp_path / "idx"])
    assert r.returncode == 0, r.stdout + r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["threw_without_index"] and info["infos"] == 1 and info["saved"] and info["opened"] and not info["wrong_window"]
    assert info["mt_ok"] and info["entries"] > 0
    out = np.fromfile(tmp_path / "out", np.uint8)
    n, ns = len(rows), info["single"]
    want = rows[:, 4].astype(np.uint8)
    assert len(out) == ns + 2 * n
    assert np.array_equal(out[:ns], want[:ns]), "single queries"
    assert np.array_equal(out[ns:ns + n], want), "bulk query"
    assert np.array_equal(out[ns + n:], want), "index reopened by a second finder"


@pytest.mark.gpu
def test_psikt_reads_a_vg_protobuf_graph(tmp_path):
    """BASELINE configs[0] names a vg graph: `psikt ... x.vg` (the protobuf stream of the reference's test/data/small/x.vg) gives
    the golden seed set of the same graph's GFA -- the .vg reader needs no protobuf library."""
    c = G["x_k12"]
    out, log = tmp_path / "seeds.bin", tmp_path / "psi.log"
    import gzip
    vg = tmp_path / "x.vg"        # the reference's test/data/small/x.vg: its protobuf stream, re-compressed
    vg.write_bytes(gzip.compress(gzip.decompress((util.GOLDEN / "inputs" / "x_vg_stream.gz").read_bytes())))
    r = run([PSIKT, "-f", util.GOLDEN / c["reads"], "-l", c["k"], "-d", c["d"], "-n", c["n_paths"], "-o", out, "-L", log, "-q", vg])
    assert r.returncode == 0, r.stderr
    g = capi.Graph.load(vg)
    got = capi.canonical(load_psikt_output(out, g))
    assert len(got) == c["count"] and util.md5_tuples(got) == c["md5"]
    assert "Input graph node IDs are" in log.read_text()
