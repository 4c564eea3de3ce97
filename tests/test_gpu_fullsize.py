"""Parity at BASELINE.json's full shapes (configs[1] chr22 shape, configs[3] MHC-like shape).

The oracle's closed form (oracle/psi_oracle.c) walks every position of the
graph, so it is run on a bounded read sample (seconds of CPU); the full read set is then
covered by size-independent properties of the seed set:
  * completeness on error-free reads: every seed (read, offset) is found at least once;
  * soundness: the locus of a sampled record spells the read's k-mer along some walk;
  * the set does not depend on chunking nor on the off-path strategy
    (materialised index entries vs. walking the graph per chunk), compared through the
    count, an order-independent checksum of the records, and sorted equality.
Bit-exact: integer / byte work.  Needs a GPU (and ~2 GB of HBM)."""
import numpy as np
import pytest

import util
from bench_support import synth
from oracle import oracle_py as orc
from psi_b200 import capi

pytestmark = pytest.mark.gpu


def build(shape):
    a = synth.graph_arrays(**synth.SHAPES[shape])
    return capi.Graph.from_arrays(a["ids"], a["seq_start"], a["seq"], a["row_ptr"], a["col"], a["path_ptr"],
                                  a["path_nodes"], sort=True)


def make_ctx(g, k, n_paths, mode=0, fused=1):
    ctx = capi.Context(k, 0)
    ctx.set_option("offpath_mode", mode)
    ctx.set_option("fused", fused)     # 0: index-mode steps through the separate seeding / probe / resolve kernels
    ctx.set_graph(g, ids="coord")
    ctx.set_paths(g.pick_paths(n_paths, seed=1))
    ctx.find_loci()
    return ctx


def run(ctx, rp, bases, d, chunk=0):
    n = len(rp) - 1
    chunk = chunk or n
    parts = []
    for b in range(0, n, chunk):
        e = min(n, b + chunk)
        ctx.submit_chunk(rp[b:e + 1] - rp[b], bases[int(rp[b]):int(rp[e])], b, d)
        cnt = ctx.seeds_all(capi.ALL)
        rec = ctx.fetch()
        assert len(rec) == cnt
        parts.append(rec)
    return np.concatenate(parts)


def sort_rows(rec):
    """records {node, node_off, read, read_off} -> rows (read, read_off, node, node_off) in canonical order"""
    t = rec.reshape(-1, 4)
    order = np.lexsort((t[:, 1], t[:, 0], t[:, 3], t[:, 2]))
    return np.ascontiguousarray(t[order][:, [2, 3, 0, 1]])


def checksum(rec):
    """order-independent: sum of a 64-bit mix of every record (mod 2^64)"""
    t = rec.reshape(-1, 4).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = t[:, 0] * np.uint64(0x9E3779B97F4A7C15)
        h ^= (t[:, 1] + np.uint64(0xBF58476D1CE4E5B9)) * np.uint64(0x94D049BB133111EB)
        h = (h ^ (h >> np.uint64(29))) * np.uint64(0xD6E8FEB86659FD93)
        h ^= (t[:, 2] + np.uint64(1)) * np.uint64(0xFF51AFD7ED558CCD)
        h = (h ^ (h >> np.uint64(32))) * np.uint64(0xC4CEB9FE1A85EC53)
        h ^= (t[:, 3] + np.uint64(7)) * np.uint64(0x2545F4914F6CDD1D)
        h ^= h >> np.uint64(31)
        return int(h.sum(dtype=np.uint64)), len(t)


def spells(g, v, o, want):
    """is there a forward walk from (node rank v, offset o) spelling `want` (bytes)?  (traverser_bfs.hpp:114-161)"""
    stack = [(v, o, 0)]
    while stack:
        v, o, i = stack.pop()
        s, e = int(g.seq_start[v]) + o, int(g.seq_start[v + 1])
        take = min(e - s, len(want) - i)
        if g.seq[s:s + take].tobytes() != want[i:i + take]:
            continue
        i += take
        if i == len(want):
            return True
        for j in range(int(g.row_ptr[v]), int(g.row_ptr[v + 1])):
            stack.append((int(g.col[j]), 0, i))
    return False


def check_sound(g, rows, rp, bases, k, n_sample, seed):
    """sampled records: the locus really spells the read's seed"""
    order = np.argsort(g.coord_id, kind="stable")
    ids_sorted = g.coord_id[order]
    rng = np.random.default_rng(seed)
    for i in rng.integers(0, len(rows), n_sample):
        r, p, nid, noff = (int(x) for x in rows[i])
        v = int(order[np.searchsorted(ids_sorted, nid)])
        assert g.coord_id[v] == nid and noff < int(g.seq_start[v + 1] - g.seq_start[v])
        want = bases[int(rp[r]) + p:int(rp[r]) + p + k].tobytes()
        assert spells(g, v, noff, want), (r, p, nid, noff)


def check_complete(rows, n_reads, read_len, k, d):
    """error-free reads: every seed of every read has at least one locus"""
    per_read = (read_len - k) // d + 1
    key = rows[:, 0] * np.uint64(per_read) + rows[:, 1] // np.uint64(d)
    seen = np.zeros(n_reads * per_read, bool)
    seen[key.astype(np.int64)] = True
    assert seen.all(), f"{(~seen).sum()} seeds of error-free reads have no locus"


@pytest.fixture(scope="module")
def chr22():
    return build("chr22")


def test_chr22_shape_sample_equals_oracle(chr22):
    """full chr22-shape graph (52 Mbp, 2.9 M nodes, 16 paths), 100 000 x 100 bp reads, k = 20: == closed form"""
    g, k = chr22, 20
    rp, bases = synth.reads(g, 100_000, 100, 77)
    want, _ = orc.seeds_closed_form(orc.OGraph.of(g), orc.OReads(rp, bases), k, k)
    for mode, fused in ((0, 1), (0, 0), (1, 1)):
        ctx = make_ctx(g, k, 16, mode, fused)
        assert ctx.counters()["offpath_mode"] == (2 if mode == 0 else 1)
        got = sort_rows(run(ctx, rp, bases, k))
        ctx.close()
        assert got.shape == want.shape and np.array_equal(got, want), f"mode {mode} fused {fused}"


def test_chr22_shape_full_read_set_properties(chr22):
    """BASELINE configs[1] at full size: 1 M x 100 bp reads, k = 20"""
    g, k, n, L = chr22, 20, 1_000_000, 100
    rp, bases = synth.reads(g, n, L, 1002)
    ctx = make_ctx(g, k, 16, 0)
    whole = run(ctx, rp, bases, k)
    rows = sort_rows(whole)
    assert not (rows[1:] == rows[:-1]).all(axis=1).any(), "duplicate records: the output must already be a set"
    check_complete(rows, n, L, k, k)
    check_sound(g, rows, rp, bases, k, 3000, 1)
    c = ctx.counters()
    assert c["n_hits"] == len(rows) and c["n_hits_on"] + c["n_hits_off"] == c["n_hits"]
    # on-path and off-path phases partition the set
    ctx.submit_chunk(rp, bases, 0, k)
    n_on = ctx.seeds_all(capi.ON_PATHS)
    cs_on = checksum(ctx.fetch())
    n_off = ctx.seeds_all(capi.OFF_PATHS)
    cs_off = checksum(ctx.fetch())
    assert n_on + n_off == len(rows)
    assert (cs_on[0] + cs_off[0]) % (1 << 64) == checksum(whole)[0]
    # chunking (ragged last chunk) does not change the set
    assert ctx.counters()["fused"] == 1
    assert checksum(run(ctx, rp, bases, k, 300_001)) == checksum(whole)
    ctx.close()
    # the separate-kernel route gives the same set as the fused kernel
    ctx = make_ctx(g, k, 16, 0, fused=0)
    unfused = run(ctx, rp, bases, k)
    assert ctx.counters()["fused"] == 0
    ctx.close()
    assert checksum(unfused) == checksum(whole)
    # walking the graph per chunk (the reference's own scheme) gives the same set
    ctx = make_ctx(g, k, 16, 1)
    walked = run(ctx, rp, bases, k, 500_000)
    assert ctx.counters()["n_walks"] > 0
    ctx.close()
    assert checksum(walked) == checksum(whole)
    assert np.array_equal(sort_rows(walked), rows)


def test_chr22_shape_150bp_reads_every_offset(chr22):
    """configs[2] read shape (150 bp) and the d = 1 stress row of SURVEY 8d on a 200 000-read shard"""
    g, k, n, L = chr22, 20, 200_000, 150
    rp, bases = synth.reads(g, n, L, 1003)
    ctx = make_ctx(g, k, 16, 0)
    rows = sort_rows(run(ctx, rp, bases, k))
    check_complete(rows, n, L, k, k)
    rows1 = sort_rows(run(ctx, rp, bases, 1, 50_000))
    check_complete(rows1, n, L, k, 1)
    check_sound(g, rows1, rp, bases, k, 2000, 2)
    # the d = k set is the sub-set of the d = 1 set at offsets divisible by k
    sub = rows1[rows1[:, 1] % np.uint64(k) == 0]
    assert np.array_equal(sub, rows)
    ctx.close()


def test_mhc_shape_k32_equals_oracle():
    """configs[3]: MHC-like hypervariable region (5 Mbp, a SNP every ~12 bp, 10 % tri-allelic), k = 32, 150 bp reads.
    Most k-walks are off the indexed paths; both off-path strategies must give the closed form."""
    g, k, n, L = build("mhc"), 32, 30_000, 150
    rp, bases = synth.reads(g, n, L, 1004)
    want, _ = orc.seeds_closed_form(orc.OGraph.of(g), orc.OReads(rp, bases), k, k)
    assert len(want) >= n * ((L - k) // k + 1)
    for mode in (1, 0):
        ctx = make_ctx(g, k, 16, mode)
        c = ctx.counters()
        got = sort_rows(run(ctx, rp, bases, k, 12_000 if mode else 0))
        assert got.shape == want.shape and np.array_equal(got, want), f"mode {mode} ({c['offpath_mode']})"
        c = ctx.counters()
        assert c["n_hits_off"] > 0
        ctx.close()


def test_chr22_shape_one_chunk_of_configs2_size(chr22):
    """BASELINE configs[2] as ONE chunk: 10 M x 150 bp reads (70 M seeds, 375 MB of 2-bit words) through the measured
    path -- 2-bit chunk, fused kernel, dense per-seed results, asynchronous step.  Completeness on the whole chunk;
    soundness of a sample; the dense planes equal what four chunks of 2.5 M give, seed for seed; and the per-hit
    records of a 1 M-read slice carry the same loci as its dense results."""
    g, k, n, L = chr22, 20, 10_000_000, 150
    per_read = (L - k) // k + 1
    parts = [synth.reads(g, 1_000_000, L, 3000 + b)[1] for b in range(10)]
    bases = np.concatenate(parts)
    del parts
    rp = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    ctx = make_ctx(g, k, 16, 0)
    ctx.submit_chunk_packed(capi.Packed.pack(rp, bases, 0), k, with_read_ptr=False)
    ctx.seeds_all_async(capi.ALL | capi.DENSE)
    cnt = ctx.wait()
    dense, extra = ctx.fetch_dense()
    assert len(dense) == n * per_read and ctx.counters()["n_seeds"] == n * per_read
    hit = dense[:, 0] != capi.NIL32
    assert hit.all(), f"{int((~hit).sum())} seeds of error-free reads found nothing"
    assert cnt == int(hit.sum()) + len(extra)
    # soundness of sampled seeds
    order = np.argsort(g.coord_id, kind="stable")
    ids_sorted = g.coord_id[order]
    rng = np.random.default_rng(11)
    for s in rng.integers(0, len(dense), 2000):
        r, j = divmod(int(s), per_read)
        v = int(order[np.searchsorted(ids_sorted, int(dense[s, 0]))])
        want = bases[r * L + j * k:r * L + j * k + k].tobytes()
        assert spells(g, v, int(dense[s, 1] & 0x7FFFFFFF), want)
    # four chunks give the same planes
    q = n // 4
    for b in range(4):
        sub = bases[b * q * L:(b + 1) * q * L]
        ctx.submit_chunk_packed(capi.Packed.pack(rp[:q + 1], sub, b * q), k, with_read_ptr=False)
        ctx.seeds_all(capi.ALL | capi.DENSE)
        d2, e2 = ctx.fetch_dense()
        assert np.array_equal(d2, dense[b * q * per_read:(b + 1) * q * per_read]), f"chunk {b}"
    # records of the first million reads against their dense results
    m = 1_000_000
    ctx.submit_chunk(rp[:m + 1], bases[:m * L], 0, k)
    ctx.seeds_all(capi.ALL)
    rows = sort_rows(ctx.fetch())
    rec, _ = capi.dense_to_records(dense[:m * per_read], extra[extra[:, 2] < m] if len(extra) else extra, rp[:m + 1], k, k, 0)
    assert np.array_equal(sort_rows(rec), rows)
    ctx.close()


def test_chr22_shape_sliced_index_build_is_the_one_shot_index(chr22):
    """The sliced index build (automatic above ~3 G estimated pairs: BASELINE configs[4]) at chr22 size: 16 slices of the
    k-mer space, paths in groups of a 2^26-window budget inside every slice -- the same number of entries, k-mers,
    starting loci and off-path entries as the one-shot build, and the same dense results (5-byte and 6-byte planes) and
    extra list for 300 000 reads, seed for seed."""
    g, k, n, L = chr22, 20, 300_000, 100
    rp, bases = synth.reads(g, n, L, 4242)
    pk = capi.Packed.pack(rp, bases, 0)
    ps = g.pick_paths(16, seed=1)
    out = []
    for slices, budget in ((1, 0), (16, 1 << 26)):
        ctx = capi.Context(k, 0)
        ctx.set_option("build_slices", slices)
        ctx.set_option("build_group_windows", budget)
        ctx.set_graph(g, ids="coord")
        ctx.set_paths(ps)
        n_loci = ctx.find_loci()
        c = ctx.counters()
        assert c["index_build_slices"] == slices
        ctx.submit_chunk_packed(pk, k, with_read_ptr=False)
        cnt = ctx.seeds_all(capi.ALL | capi.DENSE)
        dense, extra = ctx.fetch_dense()
        assert ctx.dense5_layout()[1] and ctx.seeds_all(capi.ALL | capi.DENSE5) == cnt
        dense5, extra5 = ctx.fetch_dense5()
        assert np.array_equal(dense5, dense)
        out.append((c["n_path_bases"], c["n_index_entries"], c["n_index_kmers"], c["n_offpath_entries"], n_loci, cnt, dense,
                    np.unique(extra, axis=0)))
        ctx.close()
    a, b = out
    assert a[:6] == b[:6], (a[:6], b[:6])
    assert (a[6][:, 0] != capi.NIL32).all()
    assert np.array_equal(a[6], b[6]) and np.array_equal(a[7], b[7])
