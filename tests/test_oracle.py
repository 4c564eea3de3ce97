"""The CPU restatement (oracle/psi_oracle.c) pinned against
  * the reference's own known-answer tests, and
  * golden seed sets produced by the unmodified reference (tests/golden/make_golden.py).
CPU only."""
import numpy as np
import pytest

import util
from oracle import oracle_py as orc
from psi_b200 import capi

G = util.golden_index()
CASES = {c["name"]: c for c in G["cases"]}


def load_case(c):
    g = capi.Graph.load_gfa(util.GOLDEN / c["gfa"])
    rp, bases = util.read_fasta(util.GOLDEN / c["reads"])
    return g, rp, bases


# ---- reference test/src/test_indexiter.cpp:131-402 (hit counts of kmer_exact_matches) ----

SET5_A = ["TGCAGTATAGTCGTCGCACGCCTTCTGGCCGCTGGCGGCAGTACAGGATCCTCTTGCTCACAGT"
          "GTAGGGCCCTCTTGCTCCCGGTGTGACGGCTGGCGTGCAGCTGGCTCCCCCGCTGGCAGCTGGGGACACTGACGGGCCC"
          "TCTTGCTCCCCTACTGGCCGCCTCCTGCACCAATTAAAGTCGGAGCACCGGTTACGC",
          "TGCAGTATAGTCGTCGCACGCCTTCTGGCCGCTGGCGGCAGTACAGGATCCTCTTGCTCACAGT"
          "GTAGGGCCCTCTTGCTCCCGGTGTGACGGCTGGCGTGCAGCTGGCTCCCCCGCTCGCAGGTGGCGACACAAACGGGCCC"
          "TCTTGCTCCCCTACTGGCCGCCTCCTGCACCAATTAAAGTCGGAGCACCGGTTACGC"]
SET5_B = ["CATTGCAGAGCCCTCTTGCTCACAGTGTAGTGGCAGCACGCCCGCCTCCTGGCAGCTAGGGACA"
          "GTGCCAGGCCCTCTTGCTCCAAGTGTAGTGGCAGCTGGCTCCCCCGCTGGCAGCTGGGGACACTGACGGGCCCTCTTGC"
          "TTGCAGT",
          "TAGGGCAACTGCAGGGCTATCTTGCTTACAGTGGTGTCCAGCGCCCTCTGCTGGCGTCGGAGCA"
          "TTGCAGGGCTCTCTTGCTCGCAGTGTAGTGGCGGCACGCCGCCTGCTGGCAGCTAGGGACATTGCAGAGCCCTCTTGCT"
          "CACAGTG"]


@pytest.mark.parametrize("s1,s2,k,expect", [
    (["GATAGACTAGCCA", "GGGCGTAGCCA"], ["GGGCGTAGCCA"], 4, 11),                      # test_indexiter.cpp:182-185
    (["CATATA"], ["ATATAC"], 3, 5),                                                   # :230-233
    (["TAGGCTACCGATTTAAATAGGCACAC", "TAGGCTACGGATTTAAATCGGCACAC"],
     ["GGATTTAAATA", "CGATTTAAATC", "GGATTTAAATC", "CGATTTAAATA"], 10, 8),            # :282-285
    (["TAGGCTACCGATTNAAATAGGCACAC", "TAGGCTACGGATTNAAATCGGCACAC"],
     ["GGATTNAAATA", "CGATTNAAATC", "GGATTNAAATC", "CGATTNAAATA"], 10, 0),            # :335-338 (fine iterators)
    (SET5_A, SET5_B, 30, 21),                                                         # :394-397
])
def test_kmer_exact_matches_known_answers(s1, s2, k, expect):
    assert orc.kmer_exact_matches(s1, s2, k) == expect
    assert orc.kmer_exact_matches(s2, s1, k) == expect


# ---- reference test/src/test_sequence.cpp:1194-1270 (increment_kmer) ----

def test_increment_kmer_known_answers():
    k = 20
    assert orc.increment_kmer("A" * k, k - 1) == ("A" * 19 + "C", k - 1)
    s, r = orc.increment_kmer("A" * k, 11)
    assert r == 11
    s, r = orc.increment_kmer(s, 16)
    assert (s, r) == ("AAAAAAAAAAACAAAACAAA", 16)
    assert orc.increment_kmer("A" * 10 + "T" * 10, 14) == ("AAAAAAAAACAAAAAAAAAA", 9)
    s, r = orc.increment_kmer("T" * k, k - 1)
    assert s == "A" * k and r == 2 ** 64 - 1


# ---- reference test/src/test_sequence.cpp:1272-1427 (seeding + SeedMap id/offset) ----

TRUTH_NONOVERLAP = ["CAAA", "TAAG", "AAAT", "AAGA", "TTTC", "TGGA", "ATAA", "TATT", "TTCC", "TGGT",
                    "GTCC", "TGGT", "TGCT", "ATGT", "TGTT", "GGGC", "CTTT", "TTTC", "CTTC", "TTCC"]
TRUTH_OVERLAP = ["CAAA", "AAAT", "AATA", "ATAA", "TAAG", "AAGA", "AGAT",
                 "AAAT", "AATA", "ATAA", "TAAG", "AAGA", "AGAC", "GACT",
                 "TTTC", "TTCT", "TCTG", "CTGG", "TGGA", "GGAG", "GAGT",
                 "ATAA", "TAAT", "AATA", "ATAT", "TATT", "ATTC", "TTCC",
                 "TTCC", "TCCT", "CCTG", "CTGG", "TGGT", "GGTT", "GTTG",
                 "GTCC", "TCCT", "CCTG", "CTGG", "TGGT", "GGTT", "GTTG",
                 "TGCT", "GCTA", "CTAT", "TATG", "ATGT", "TGTG", "GTGT",
                 "TGTT", "GTTG", "TTGG", "TGGG", "GGGC", "GGCT", "GCTT",
                 "CTTT", "TTTT", "TTTT", "TTTT", "TTTC", "TTCT", "TCTT",
                 "CTTC", "TTCT", "TCTT", "CTTC", "TTCC", "TCCT", "CCTT"]


@pytest.mark.parametrize("d,truth,per_read", [(4, TRUTH_NONOVERLAP, 2), (1, TRUTH_OVERLAP, 7)])
def test_seeding_known_answers(d, truth, per_read):
    rp, bases = util.read_fasta(util.GOLDEN / "inputs/reads_n10l10e0i0.fa")
    reads = orc.OReads(rp, bases)
    rid, off = orc.seeding(reads, 4, d)
    assert len(rid) == len(truth)
    for i, t in enumerate(truth):
        s = int(rp[int(rid[i])]) + int(off[i])
        assert bases[s:s + 4].tobytes().decode() == t
        assert rid[i] == i // per_read                      # position_to_id
        assert off[i] == (i % per_read) * d                 # position_to_offset of the seed's first base


def test_seeding_short_reads_and_offset():
    rp = np.array([0, 3, 13, 13, 20], np.uint64)
    bases = np.frombuffer(b"ACG" + b"ACGTACGTAC" + b"" + b"ACGTACG", np.uint8)
    rid, off = orc.seeding(orc.OReads(rp, bases, first_read_id=100), 4, 4)
    assert rid.tolist() == [101, 101, 103] and off.tolist() == [0, 4, 0]


# ---- reference test/src/test_traverser.cpp:81-96 ----

def test_traverser_known_answers():
    g = capi.Graph.load_gfa(util.GOLDEN / "inputs/x.gfa.gz")
    rp, bases = util.read_fasta(util.GOLDEN / "inputs/reads_n10l10e0i0.fa")
    node, off = util.all_loci(g)
    t, raw = orc.seeds_off_paths(orc.OGraph.of(g), node, off, orc.OReads(rp, bases), 10, 10)
    truth = [(1, 0), (1, 1), (9, 4), (9, 17), (16, 0), (17, 0), (20, 0), (20, 31), (20, 38), (20, 38)]
    assert t.tolist() == [[i, 0, n, o] for i, (n, o) in enumerate(truth)]


# ---- loader parity with gum ----

@pytest.mark.parametrize("name", ["tiny", "x", "multi", "m"])
def test_loader_reproduces_gum_ranks_and_ids(name):
    ref = np.load(util.GOLDEN / f"nodes_{name}.npy")
    g = capi.Graph.load_gfa(util.GOLDEN / f"inputs/{name}.gfa.gz")
    assert np.array_equal(ref[:, 0], g.internal_id)
    assert np.array_equal(ref[:, 1], g.coord_id)
    assert np.array_equal(ref[:, 2], g.seq_start[1:] - g.seq_start[:-1])
    # closed form of SURVEY 8a-8
    deg = (g.row_ptr[1:] - g.row_ptr[:-1]).astype(np.int64)
    indeg = np.bincount(g.col, minlength=g.n_nodes).astype(np.int64)
    ids = 1 + np.concatenate([[0], np.cumsum(5 + 3 * (deg + indeg))[:-1]])
    assert np.array_equal(ids.astype(np.uint64), g.internal_id)


# ---- golden seed sets of the compiled reference ----

@pytest.mark.parametrize("name", sorted(CASES))
def test_closed_form_matches_reference_golden(name):
    c = CASES[name]
    if c["query_seeds"] > 200000:
        pytest.skip("large case: covered on the GPU box and by test_big_golden_cpu")
    g, rp, bases = load_case(c)
    t, raw = orc.seeds_closed_form(orc.OGraph.of(g), orc.OReads(rp, bases), c["k"], c["d"])
    assert len(t) == c["count"]
    assert util.md5_tuples(t) == c["md5"]


def test_big_golden_cpu():
    c = CASES["x_k20_d1"]
    g, rp, bases = load_case(c)
    t, raw = orc.seeds_closed_form(orc.OGraph.of(g), orc.OReads(rp, bases), c["k"], c["d"])
    assert len(t) == c["count"] and util.md5_tuples(t) == c["md5"]


@pytest.mark.parametrize("name", ["x_k12", "m_k20", "fuzz_03", "fuzz_06", "fuzz_10", "multi_k32"])
def test_hybrid_decomposition_equals_closed_form(name):
    """seeds_on_paths(paths) U seeds_off_paths(uncovered loci) == closed form, with this
    repo's own (seeded) path picker: the set does not depend on the paths."""
    c = CASES[name]
    g, rp, bases = load_case(c)
    og, orr = orc.OGraph.of(g), orc.OReads(rp, bases)
    ps = g.pick_paths(c["n_paths"], seed=7)
    op = orc.OPaths(ps.path_ptr, ps.nodes, ps.head_off, ps.tail_trim)
    ln, lo = orc.uncovered_loci(og, op, c["k"])
    t_all, _ = orc.seeds_all(og, op, ln, lo, orr, c["k"], c["d"])
    assert util.md5_tuples(t_all) == c["md5"]
    t_on, _ = orc.seeds_on_paths(og, op, orr, c["k"], c["d"])
    t_off, _ = orc.seeds_off_paths(og, ln, lo, orr, c["k"], c["d"])
    union = np.unique(np.concatenate([t_on, t_off]), axis=0)
    assert util.md5_tuples(union) == c["md5"]
    # removing any uncovered locus loses nothing only if it carried no read hit; all loci are needed for coverage:
    n_all, _ = util.all_loci(g)
    assert len(ln) <= len(n_all)


def test_chunking_does_not_change_the_set():
    c = CASES["x_k12"]
    g, rp, bases = load_case(c)
    og = orc.OGraph.of(g)
    parts = []
    n = len(rp) - 1
    for b in range(0, n, 2500):
        e = min(n, b + 2500)
        sub_ptr = rp[b:e + 1] - rp[b]
        sub_bases = bases[int(rp[b]):int(rp[e])]
        t, _ = orc.seeds_closed_form(og, orc.OReads(sub_ptr, sub_bases, first_read_id=b), c["k"], c["d"])
        parts.append(t)
    t = np.concatenate(parts)
    assert util.md5_tuples(t) == c["md5"]   # disjoint read ranges: concatenation is already canonical


# ---- starting loci pinned to the reference (tests/golden/make_loci_golden.py) ----
#
# The reference derives its starting loci from NODE-level coverage: a k-walk is covered when the sequence of nodes it
# runs through is a contiguous piece of some indexed path (add_uncovered_loci + covered_by, reference
# seed_finder.hpp:1481-1541, pathset.hpp:205-218,389-395).  This build (oracle psi_oracle_uncovered_loci = the device's
# find_loci_kernel, compared with each other in tests/test_gpu_parity.py) uses the k-MER level: a locus starts a walk
# iff some k-walk from it spells a k-mer w with (w, locus) absent from the path index.  On the same paths the two
# relate as: ours is a subset of the reference's, and every locus only the reference has is redundant -- all its
# k-walks are already on-path entries (the walk leaves the path's nodes but spells what the path spells there).

import glob  # noqa: E402
import os  # noqa: E402

LOCI_FIXTURES = sorted(glob.glob(os.fspath(util.GOLDEN / "loci" / "*.npz")))


def _load_loci_fixture(path):
    z = np.load(path)
    g = capi.Graph.load_gfa(util.GOLDEN / str(z["gfa"]))
    return z, g, orc.OPaths(z["path_ptr"], z["nodes"], z["head"], z["tail"])


def _path_pairs(g, z, k):
    """{(k-mer bytes, global position)} of all k-windows of the fixture's paths."""
    pairs = set()
    for p in range(len(z["head"])):
        nodes = z["nodes"][int(z["path_ptr"][p]):int(z["path_ptr"][p + 1])]
        seq = np.concatenate([g.seq[int(g.seq_start[v]):int(g.seq_start[v + 1])] for v in nodes])
        gpos = np.concatenate([np.arange(int(g.seq_start[v]), int(g.seq_start[v + 1])) for v in nodes])
        lo, hi = int(z["head"][p]), len(seq) - int(z["tail"][p])
        text, gp = seq[lo:hi].tobytes().upper(), gpos[lo:hi]
        for i in range(len(text) - k + 1):
            w = text[i:i + k]
            if set(w) <= set(b"ACGT"):
                pairs.add((w, int(gp[i])))
    return pairs


def _walks_from(g, v, off, k):
    """All k-mers spelt by forward walks from (node rank v, offset off); N stops a walk (traverser_bfs.hpp:114-161)."""
    out, stack = [], [(v, off, b"")]
    while stack:
        v, o, acc = stack.pop()
        lab = g.seq[int(g.seq_start[v]) + o:int(g.seq_start[v + 1])].tobytes().upper()
        acc += lab[:k - len(acc)]
        if not set(acc) <= set(b"ACGT"):
            continue
        if len(acc) == k:
            out.append(acc)
            continue
        for e in range(int(g.row_ptr[v]), int(g.row_ptr[v + 1])):
            stack.append((int(g.col[e]), 0, acc))
    return out


@pytest.mark.parametrize("fixture", [f for f in LOCI_FIXTURES if "_e" not in os.path.basename(f)], ids=lambda f: os.path.basename(f)[:-4])
def test_starting_loci_against_the_reference_on_its_own_paths(fixture):
    z, g, op = _load_loci_fixture(fixture)
    k = int(z["k"])
    n, o = orc.uncovered_loci(orc.OGraph.of(g), op, k)
    ours = set(zip(n.tolist(), o.tolist()))
    ref = set(zip(z["loci_rank"].tolist(), z["loci_off"].tolist()))
    assert ours <= ref, sorted(ours - ref)[:5]
    extra = ref - ours
    assert len(extra) <= 0.02 * max(len(ref), 1) + 4, "the two definitions differ on a handful of loci only"
    if extra and g.n_nodes < 1000:
        pairs = _path_pairs(g, z, k)
        for v, off in extra:
            gp = int(g.seq_start[v]) + off
            for w in _walks_from(g, v, off, k):
                assert (w, gp) in pairs, "a locus only the reference starts from must be fully served by the path index"


def test_reference_known_answer_for_the_tiny_graph():
    """reference test/src/test_seedfinder.cpp:102-126: tiny graph, k = 12, 4 paths -> loci (1, 2..7), (2, 0), (3, 0);
    8 (and 32) paths -> none.  The fixtures were produced by the reference itself and carry exactly that."""
    z, g, op = _load_loci_fixture(util.GOLDEN / "loci" / "tiny_k12_n4.npz")
    got = sorted((int(g.coord_id[v]), int(o)) for v, o in zip(z["loci_rank"], z["loci_off"]))
    assert got == [(1, 2), (1, 3), (1, 4), (1, 5), (1, 6), (1, 7), (2, 0), (3, 0)]
    n, o = orc.uncovered_loci(orc.OGraph.of(g), op, 12)
    assert sorted((int(g.coord_id[v]), int(x)) for v, x in zip(n, o)) == got
    z8, g8, op8 = _load_loci_fixture(util.GOLDEN / "loci" / "tiny_k12_n8.npz")
    assert len(z8["loci_rank"]) == 0 and len(orc.uncovered_loci(orc.OGraph.of(g8), op8, 12)[0]) == 0


# ---- MEM mode pinned to the reference (tests/golden/make_mem_golden.py) ----

MEM_FIXTURES = sorted(glob.glob(os.fspath(util.GOLDEN / "mems" / "*.npz")))


def load_mem_fixture(path):
    z = np.load(path)
    g = capi.Graph.load_gfa(util.GOLDEN / str(z["gfa"]))
    rp, bases = util.read_fasta(util.GOLDEN / str(z["reads"]))
    n = min(int(z["max_reads"]), len(rp) - 1)
    return z, g, rp[:n + 1], bases[:int(rp[n])]


@pytest.mark.parametrize("fixture", MEM_FIXTURES, ids=lambda f: os.path.basename(f)[:-4])
def test_find_mems_restatement_against_the_reference(fixture):
    """oracle_py.find_mems (restatement of index_iter.hpp:854-906) on the reference's own paths gives the hits the
    reference's seeds_on_paths(sequence, cb) gave: offsets, lengths (a gocc threshold extends them, up to 63 here),
    occurrence counts, loci; max_mem cuts a read off after that many raw hits."""
    z, g, rp, bases = load_mem_fixture(fixture)
    texts, gpos = orc.path_texts(g, z["path_ptr"], z["nodes"], z["head"], z["tail"])
    rows = []
    n_check = min(len(rp) - 1, 100)          # pure-Python substring searches: a sample of the reads keeps the suite short
    for r in range(n_check):
        pat = bases[int(rp[r]):int(rp[r + 1])].tobytes()
        for st, pl, go, ti, o in orc.find_mems(texts, pat, int(z["k"]), int(z["gocc"]), int(z["max_mem"])):
            gp = int(gpos[ti][o])
            v = int(np.searchsorted(g.seq_start, gp, side="right") - 1)
            rows.append((r, st, pl, go, int(g.coord_id[v]), gp - int(g.seq_start[v])))
    got = np.unique(np.array(rows, np.uint64).reshape(-1, 6), axis=0)
    want = z["mems"][z["mems"][:, 0] < n_check]
    assert len(want) > 50 and np.array_equal(got, want)


# ------------------------------------------------------------------------------------------------------------------
# Paired-end distance verification: the restatement against the unmodified reference (tests/golden/make_dist_golden.py)

DIST_FIXTURES = sorted(glob.glob(os.fspath(util.GOLDEN / "dist" / "*.npz")))


def test_distance_fixtures_exist():
    assert len(DIST_FIXTURES) >= 8


@pytest.mark.parametrize("fixture", DIST_FIXTURES, ids=lambda f: os.path.basename(f)[:-4])
def test_verify_distance_restatement_against_the_reference(fixture):
    """oracle_py.verify_distance (walks of dmin..dmax characters between two loci) answers every locus pair like
    SeedFinder::verify_distance over DiVerG's matrix index (seed_finder.hpp:1193-1317) did in the compiled reference."""
    z = np.load(fixture)
    og = orc.OGraph.of(capi.Graph.load_gfa(util.GOLDEN / str(z["gfa"])))
    dmin, dmax = int(z["dmin"]), int(z["dmax"])
    rows = z["rows"]
    assert rows[:, 4].any() and not rows[:, 4].all()
    got = np.array([orc.verify_distance(og, int(v), int(o), int(u), int(p), dmin, dmax) for v, o, u, p, _ in rows])
    assert np.array_equal(got, rows[:, 4].astype(bool))


def test_verify_distance_known_answers_of_the_reference_test():
    """test/src/test_seedfinder.cpp:225-268: tiny graph, window 8..12, the pairs the reference lists as distant / closed."""
    g = capi.Graph.load_gfa(util.GOLDEN / "inputs/tiny.gfa.gz")
    og = orc.OGraph.of(g)
    r = {int(c): i for i, c in enumerate(g.coord_id)}
    distant = [(1, 0, 1, 0), (1, 0, 1, 1), (1, 0, 1, 3), (1, 0, 1, 6), (1, 0, 1, 7), (1, 0, 7, 0), (2, 0, 9, 10), (9, 1, 9, 14),
               (9, 5, 9, 18), (9, 18, 11, 0), (9, 18, 11, 3), (9, 18, 15, 0), (9, 18, 15, 6)]
    closed = [(1, 0, 2, 0), (1, 0, 6, 0), (1, 0, 6, 2), (9, 0, 9, 8), (9, 1, 9, 13), (9, 10, 9, 18), (9, 6, 9, 18), (9, 18, 15, 1),
              (9, 18, 15, 5)]
    for a, b, c, d in distant:
        assert not orc.verify_distance(og, r[a], b, r[c], d, 8, 12)
    for a, b, c, d in closed:
        assert orc.verify_distance(og, r[a], b, r[c], d, 8, 12)
