"""Host-side logic and the C-ABI surface, CPU only (no compute calls)."""
import ctypes as C
import gzip
import os
import re

import numpy as np
import pytest

import util
from psi_b200 import capi


def test_library_exports_every_declared_symbol():
    header = (util.ROOT / "include" / "psi_b200.h").read_text()
    declared = set(re.findall(r"\b(psi_b200_[a-z0-9_]+)\s*\(", header))
    declared -= {"psi_b200_graph", "psi_b200_ctx"}
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    L = capi.lib()
    for s in declared:
        assert hasattr(L, s), f"{s} is declared in include/psi_b200.h but not exported"
    assert L.psi_b200_version().decode().endswith("sm_100a")


def test_host_library_has_the_host_half_only_and_maps_no_cuda():
    """libpsi_b200_host.so: graphs, paths and reads without the CUDA runtime (what bench.py's reference arm binds)."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from psi_b200 import capi\n"
        "capi.use_host_library()\n"
        "L = capi.lib()\n"
        "assert all(hasattr(L, s) for s in capi.HOST_SYMBOLS)\n"
        "assert not hasattr(L, 'psi_b200_create') and not hasattr(L, 'psi_b200_seeds_all')\n"
        "g = capi.Graph.load_gfa(%r)\n"
        "ps = g.pick_paths(2)\n"
        "assert g.n_nodes == 210 and ps.n_paths == 2\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'libpsi_b200_host.so' in maps and 'libpsi_b200.so' not in maps and 'libcudart' not in maps, 'CUDA mapped'\n"
        "print('ok')\n") % (os.fspath(util.ROOT), os.fspath(util.GOLDEN / "inputs/x.gfa.gz"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr[-1000:]


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.PsiError) as e:
        capi.Context(12, 0)
    assert e.value.code == capi.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    for p in (util.ROOT / "psi_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".hpp", ".h") and p.is_file():
            txt = p.read_text()
            assert "oracle" not in txt.lower(), f"{p} mentions the oracle"


def test_reader_fasta_fastq_gz_and_chunking(tmp_path):
    fq = tmp_path / "r.fastq"
    fq.write_text("@a desc\nACGT\n+\nIIII\n@b\nGGCCA\n+\nIIIII\n@c\nTT\n+\nII\n")
    r = capi.Reader(fq)
    first, rp, bases, names = r.next(2)
    assert (first, rp.tolist(), bases.tobytes(), names) == (0, [0, 4, 9], b"ACGTGGCCA", ["a", "b"])
    first, rp, bases, names = r.next(2)
    assert (first, rp.tolist(), bases.tobytes(), names) == (2, [0, 2], b"TT", ["c"])   # global ids (sequence.hpp:1616)
    assert r.next(2) is None
    fa = tmp_path / "r.fa.gz"
    with gzip.open(fa, "wt") as f:
        f.write(">x\nACGT\nACG\n>y\nTTTT\n")
    first, rp, bases, names = capi.Reader(fa).next(0)
    assert (rp.tolist(), bases.tobytes(), names) == ([0, 7, 11], b"ACGTACGTTTT", ["x", "y"])
    with pytest.raises(capi.PsiError) as e:
        capi.Reader(tmp_path / "missing.fq")
    assert e.value.code == capi.ERR_IO


def _unpack(words, n):
    w = np.asarray(words, np.uint64)
    idx = np.arange(n)
    return ((w[idx // 32] >> ((idx % 32) * 2).astype(np.uint64)) & np.uint64(3)).astype(np.uint8)


def test_pack_bases_matches_a_plain_restatement():
    """psi_b200_pack_bases (byte-parallel packer of the chunk reader) against numpy: codes A,C,G,T = 0..3 of either
    case, everything else code 0 + its position in the exception list; every length around the 8- and 32-base groups."""
    rng = np.random.default_rng(1)
    alphabet = np.frombuffer(b"ACGTacgtNn-*RYKM", np.uint8)
    code = np.zeros(256, np.uint8)
    for ch, v in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
        code[ch] = v
    valid = np.zeros(256, bool)
    valid[list(b"ACGTacgt")] = True
    for n in list(range(0, 70)) + [127, 128, 129, 1000, 4097]:
        b = alphabet[rng.choice(len(alphabet), n, p=[0.2] * 4 + [0.02] * 4 + [0.015] * 8)]
        rp = np.array([0, n], np.uint64)
        pk = capi.Packed.pack(rp, b)
        assert len(pk.words) == n // 32 + 2
        assert np.array_equal(_unpack(pk.words, n), code[b]), n
        assert np.array_equal(pk.exc, np.flatnonzero(~valid[b]).astype(np.uint64)), n
        tail = np.asarray(pk.words, np.uint64)
        assert n % 32 == 0 or int(tail[n // 32]) >> (2 * (n % 32)) == 0      # zero filled past the last base
        assert int(tail[-1]) == 0


def test_reader_delivers_packed_chunks_and_alternates_its_buffers(tmp_path):
    """psi_b200_reader_next_packed: the same records as the character reader, as 2-bit words; equal-length chunks
    say so; what a call returned stays intact while the next chunk is parsed (two buffer sets alternate)."""
    fq = tmp_path / "r.fastq"
    reads = ["ACGTACGTAC", "GGNCCATTAG", "TTTTTTTTTT", "acgtnACGTA", "CCCCCCCCCC", "ACG"]
    fq.write_text("".join(f"@r{i}\n{s}\n+\n{'I' * len(s)}\n" for i, s in enumerate(reads)))
    r = capi.Reader(fq)
    a = r.next_packed(2)
    ref = capi.Packed.pack(np.array([0, 10, 20], np.uint64), np.frombuffer("".join(reads[:2]).encode(), np.uint8))
    assert a.read_len == 10 and a.first_read_id == 0 and np.array_equal(a.read_ptr, ref.read_ptr)
    assert np.array_equal(a.words, ref.words) and a.exc.tolist() == [12]
    # raw view of the first chunk, then parse the second: the first chunk's buffers must not change
    v1 = capi.PackedChunk()
    r2 = capi.Reader(fq)
    capi._check(capi.lib().psi_b200_reader_next_packed(r2._h, 2, C.byref(v1)))
    w1 = np.ctypeslib.as_array(C.cast(v1.words, C.POINTER(C.c_uint64)), shape=(2,)).copy()
    v2 = capi.PackedChunk()
    capi._check(capi.lib().psi_b200_reader_next_packed(r2._h, 2, C.byref(v2)))
    assert v2.first_read_id == 2 and v2.words != v1.words
    assert np.array_equal(np.ctypeslib.as_array(C.cast(v1.words, C.POINTER(C.c_uint64)), shape=(2,)), w1)
    b = r.next_packed(2)
    assert b.first_read_id == 2 and b.read_len == 10 and b.exc.tolist() == [14]
    c = r.next_packed(0)
    assert c.first_read_id == 4 and c.read_len == 0 and c.read_ptr.tolist() == [0, 10, 13]     # ragged: offsets needed
    assert r.next_packed(2) is None


def test_gfa1_and_gfa2_load_the_same_graph(tmp_path):
    g2 = capi.Graph.load_gfa(util.GOLDEN / "inputs/x.gfa.gz")
    p = tmp_path / "x1.gfa"
    g2.write_gfa(p)
    g1 = capi.Graph.load_gfa(p)
    for a in ("seq_start", "seq", "row_ptr", "col", "internal_id", "coord_id"):
        assert np.array_equal(getattr(g1, a), getattr(g2, a)), a
    assert g1.path(0)[1].tolist() == g2.path(0)[1].tolist()


def test_from_arrays_matches_loader():
    g = capi.Graph.load_gfa(util.GOLDEN / "inputs/multi.gfa.gz")
    # shuffle the nodes, rebuild from arrays with sort=True: same graph
    perm = np.random.default_rng(3).permutation(g.n_nodes)
    inv = np.argsort(perm)
    lens = (g.seq_start[1:] - g.seq_start[:-1])[perm]
    seq_start = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    seq = np.concatenate([g.seq[int(g.seq_start[v]):int(g.seq_start[v + 1])] for v in perm])
    deg = (g.row_ptr[1:] - g.row_ptr[:-1])[perm]
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint64)
    col = np.concatenate([inv[g.col[int(g.row_ptr[v]):int(g.row_ptr[v + 1])]] for v in perm]).astype(np.uint32)
    paths = [g.path(i)[1] for i in range(g.n_paths)]
    path_ptr = np.concatenate([[0], np.cumsum([len(p) for p in paths])]).astype(np.uint64)
    path_nodes = np.concatenate([inv[p] for p in paths]).astype(np.uint32)
    h = capi.Graph.from_arrays(g.coord_id[perm], seq_start, seq, row_ptr, col, path_ptr, path_nodes, sort=True)
    for a in ("seq_start", "seq", "row_ptr", "col", "internal_id", "coord_id"):
        assert np.array_equal(getattr(h, a), getattr(g, a)), a


@pytest.mark.parametrize("name,n", [("tiny", 4), ("x", 16), ("multi", 4), ("m", 8)])
def test_pick_paths_are_walks_from_each_region_start(name, n):
    g = capi.Graph.load_gfa(util.GOLDEN / f"inputs/{name}.gfa.gz")
    ps = g.pick_paths(n, seed=11)
    assert ps.n_paths == n * g.n_paths
    succ = [set(g.col[int(g.row_ptr[v]):int(g.row_ptr[v + 1])].tolist()) for v in range(g.n_nodes)]
    for i in range(ps.n_paths):
        nodes = ps.nodes[int(ps.path_ptr[i]):int(ps.path_ptr[i + 1])]
        region = g.path(i // n)[1]
        assert nodes[0] == region[0]
        assert all(int(b) in succ[int(a)] for a, b in zip(nodes[:-1], nodes[1:]))
        assert len(succ[int(nodes[-1])]) == 0           # ends in a sink
    # seeded: reproducible; different seeds give different haplotypes on a variant-dense graph
    ps2 = g.pick_paths(n, seed=11)
    assert np.array_equal(ps.nodes, ps2.nodes)
    if name == "m":
        assert not np.array_equal(ps.nodes, g.pick_paths(n, seed=12).nodes)


def test_pick_paths_requires_an_embedded_path(tmp_path):
    p = tmp_path / "nopath.gfa"
    p.write_text("S\t1\tACGT\nS\t2\tGG\nL\t1\t+\t2\t+\t0M\n")
    g = capi.Graph.load_gfa(p)
    with pytest.raises(capi.PsiError) as e:       # reference seed_finder.hpp:1145-1147 throws
        g.pick_paths(2)
    assert e.value.code == capi.ERR_ARG


@pytest.mark.parametrize("name", ["x_k12_n16", "x_k20_n8", "multi_k32_n4", "m_k20_n4", "fuzz02_k12_n3_full"])
def test_paths_of_an_index_saved_by_the_reference_are_read_back(name):
    """psi_b200_pathset_load_reference decodes `<prefix>_paths` as the unmodified reference wrote it
    (tests/golden/make_refindex_golden.py): sdsl enc_vector<elias_delta, 128> of coordinate ids (samples + delta codes,
    paths of up to 3 400 nodes), left / right trims -> the same ranks, head offsets and tail trims the driver dumped."""
    z = np.load(util.GOLDEN / "refindex" / f"{name}.npz")
    g = capi.Graph.load_gfa(util.GOLDEN / str(z["gfa"]))
    ps, context = g.load_reference_paths(util.GOLDEN / "refindex" / f"{name}_paths")
    assert context == (int(z["k"]) if bool(z["patched"]) else 0)
    assert np.array_equal(ps.path_ptr, z["path_ptr"]) and np.array_equal(ps.nodes, z["nodes"])
    assert np.array_equal(ps.head_off, z["head"]) and np.array_equal(ps.tail_trim, z["tail"])
    # a file for another graph, a truncated file, a missing file: refused
    other = capi.Graph.load_gfa(util.GOLDEN / "inputs/tiny.gfa.gz")
    with pytest.raises(capi.PsiError) as e:
        other.load_reference_paths(util.GOLDEN / "refindex" / f"{name}_paths")
    assert e.value.code == capi.ERR_IO
    with pytest.raises(capi.PsiError):
        g.load_reference_paths(util.GOLDEN / "refindex" / "no_such_paths")


def test_truncated_reference_paths_file_is_refused(tmp_path):
    raw = (util.GOLDEN / "refindex" / "x_k20_n8_paths").read_bytes()
    g = capi.Graph.load_gfa(util.GOLDEN / "inputs/x.gfa.gz")
    for cut in (5, 24, 100, 700):                 # inside the paths section (what follows it is not read)
        p = tmp_path / "cut_paths"
        p.write_bytes(raw[:cut])
        with pytest.raises(capi.PsiError) as e:
            g.load_reference_paths(p)
        assert e.value.code == capi.ERR_IO


def test_reader_line_views_across_buffer_refills_and_threaded_packing(tmp_path):
    """The reader parses lines as views into a 4 MB read buffer and packs large chunks over several threads: records
    that straddle refills, a sequence line longer than the buffer, CRLF line ends, blank lines, multi-line FASTA, FASTQ
    whose quality line starts with '@' or '+', no newline at the end of the file -- against a plain Python parser; the
    threaded packer against the single-threaded one."""
    rng = np.random.default_rng(9)
    acgt = np.frombuffer(b"ACGTN", np.uint8)

    def seq(n):
        return acgt[rng.choice(5, n, p=[0.2475] * 4 + [0.01])].tobytes()
    # FASTA: multi-line records of many lengths, one 9 MB line, CRLF here and there, blank lines between records
    recs = [(b"r%d" % i, seq(int(rng.integers(1, 400)))) for i in range(30000)]
    recs.insert(777, (b"long", seq(9_000_000)))
    fa = tmp_path / "big.fa"
    with open(fa, "wb") as f:
        for i, (name, s) in enumerate(recs):
            eol = b"\r\n" if i % 7 == 0 else b"\n"
            f.write(b">" + name + b" some description" + eol)
            if name == b"long" or i % 3:
                f.write(s + eol)
            else:                                    # wrapped at 60 columns
                for j in range(0, len(s), 60):
                    f.write(s[j:j + 60] + eol)
            if i % 11 == 0:
                f.write(eol)
        f.write(b">last\nACGT")                      # no newline at the end
    recs.append((b"last", b"ACGT"))
    r = capi.Reader(fa)
    got_names, got = [], []
    while True:
        ch = r.next(5000)
        if ch is None:
            break
        first, rp, bases, names = ch
        assert first == len(got)
        got_names += names
        got += [bases[int(rp[i]):int(rp[i + 1])].tobytes() for i in range(len(rp) - 1)]
    assert got_names == [n.decode() for n, _ in recs]
    assert got == [s for _, s in recs]
    # the same file as packed chunks: words and exceptions equal to packing the characters in one thread
    r = capi.Reader(fa)
    at = 0
    while True:
        pk = r.next_packed(20000)
        if pk is None:
            break
        n = pk.n_reads
        flat = np.frombuffer(b"".join(s for _, s in recs[at:at + n]), np.uint8)
        ref = capi.Packed.pack(pk.read_ptr, flat, at)
        assert np.array_equal(pk.words, ref.words) and np.array_equal(pk.exc, ref.exc) and pk.first_read_id == at
        at += n
    assert at == len(recs)
    # FASTQ whose quality strings begin with the record markers
    fq = tmp_path / "q.fastq"
    fq.write_bytes(b"@a\nACGT\n+a\n@III\n@b x\nGG\n+\n+I\n@c\nT\n+\nI")
    first, rp, bases, names = capi.Reader(fq).next(0)
    assert (rp.tolist(), bases.tobytes(), names) == ([0, 4, 6, 7], b"ACGTGGT", ["a", "b", "c"])


# ------------------------------------------------------------------------------------------------------------------
# vg protobuf graphs (psi_b200_graph_load_vg / psi_b200_graph_load)

def _pb_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _pb_field(num, wire, payload):
    return _pb_varint((num << 3) | wire) + (payload if wire == 0 else _pb_varint(len(payload)) + payload)


def _vg_node(seq, nid):
    return _pb_field(1, 2, seq.encode()) + _pb_field(2, 2, b"n%d" % nid) + _pb_field(3, 0, _pb_varint(nid))


def _vg_edge(a, b, from_start=False):
    e = _pb_field(1, 0, _pb_varint(a)) + _pb_field(2, 0, _pb_varint(b))
    if from_start:
        e += _pb_field(3, 0, _pb_varint(1))
    return e + _pb_field(5, 0, _pb_varint(0))


def _vg_mapping(nid, rank):
    pos = _pb_field(1, 0, _pb_varint(nid)) + _pb_field(2, 0, _pb_varint(0))
    edit = _pb_field(1, 0, _pb_varint(3)) + _pb_field(2, 0, _pb_varint(3))
    return _pb_field(1, 2, pos) + _pb_field(2, 2, edit) + _pb_field(5, 0, _pb_varint(rank))


def _vg_stream(graphs, tagged=True):
    out = bytearray()
    for gmsg in graphs:
        msgs = ([b"VG"] if tagged else []) + [gmsg]
        out += _pb_varint(len(msgs))
        for m in msgs:
            out += _pb_varint(len(m)) + m
    return bytes(out)


def vg_fixture(name, tmp_path, bgzf=True):
    """tests/golden/inputs/<name>_vg_stream.gz holds the protobuf stream of the reference's test/data/*/<name>.vg (the
    file's BGZF container re-encoded); written out as a .vg file again -- as concatenated gzip members, the way BGZF
    stores it, or uncompressed."""
    raw = gzip.decompress((util.GOLDEN / "inputs" / f"{name}_vg_stream.gz").read_bytes())
    f = tmp_path / f"{name}.vg"
    f.write_bytes(b"".join(gzip.compress(raw[i:i + 200]) for i in range(0, len(raw), 200)) if bgzf else raw)
    return f


@pytest.mark.parametrize("name", ["tiny", "x", "multi"])
def test_vg_file_loads_the_same_graph_as_its_gfa(name, tmp_path):
    """The reference's own test graphs exist as .vg and as .gfa: both readers must give the same flattened graph -- ranks,
    ids, labels, out-edge order, embedded path -- and psi_b200_graph_load chooses by file name like gum::util::load."""
    a = capi.Graph.load(vg_fixture(name, tmp_path))
    assert np.array_equal(capi.Graph.load(vg_fixture(name, tmp_path, bgzf=False)).col, a.col)
    b = capi.Graph.load_gfa(util.GOLDEN / "inputs" / f"{name}.gfa.gz")
    for attr in ("seq_start", "seq", "row_ptr", "col", "coord_id", "internal_id"):
        assert np.array_equal(getattr(a, attr), getattr(b, attr)), attr
    assert a.n_paths == b.n_paths >= 1
    for i in range(a.n_paths):
        assert a.path(i)[0] == b.path(i)[0] and np.array_equal(a.path(i)[1], b.path(i)[1])
    c = capi.Graph.load_vg(vg_fixture(name, tmp_path))
    assert np.array_equal(c.col, a.col)


def test_vg_reader_chunks_untagged_streams_unknown_fields_and_errors(tmp_path):
    """Hand-encoded vg streams: a graph split over two chunks (edges before their nodes' chunk, a path continued in the
    second chunk with its mappings out of rank order), with and without type tags, raw and gzip-compressed, with fields
    this reader does not know; damaged and foreign files are refused."""
    nodes = {1: "ACGT", 2: "G", 3: "T", 4: "CCA"}
    chunk1 = b"".join(_pb_field(1, 2, _vg_node(nodes[i], i)) for i in (3, 1)) + _pb_field(2, 2, _vg_edge(1, 2)) + \
        _pb_field(2, 2, _vg_edge(1, 3)) + _pb_field(3, 2, _pb_field(2, 2, _vg_mapping(2, 2)) + _pb_field(2, 2, _vg_mapping(1, 1)) +
                                                    _pb_field(1, 2, b"ref")) + _pb_field(9, 0, _pb_varint(77))
    chunk2 = b"".join(_pb_field(1, 2, _vg_node(nodes[i], i)) for i in (4, 2)) + _pb_field(2, 2, _vg_edge(2, 4)) + \
        _pb_field(2, 2, _vg_edge(3, 4)) + _pb_field(3, 2, _pb_field(1, 2, b"ref") + _pb_field(2, 2, _vg_mapping(4, 3)) +
                                                    _pb_field(7, 2, b"ignored"))
    gfa = tmp_path / "ref.gfa"
    gfa.write_text("S\t1\tACGT\nS\t2\tG\nS\t3\tT\nS\t4\tCCA\nL\t1\t+\t2\t+\t0M\nL\t1\t+\t3\t+\t0M\nL\t2\t+\t4\t+\t0M\nL\t3\t+\t4\t+\t0M\n"
                   "P\tref\t1+,2+,4+\t*\n")
    want = capi.Graph.load_gfa(gfa)
    for tagged in (True, False):
        for compress in (False, True):
            raw = _vg_stream([chunk1, chunk2], tagged)
            f = tmp_path / f"g_{int(tagged)}{int(compress)}.vg"
            f.write_bytes(gzip.compress(raw) if compress else raw)
            g = capi.Graph.load(f)
            for attr in ("seq_start", "seq", "row_ptr", "col", "coord_id"):
                assert np.array_equal(getattr(g, attr), getattr(want, attr)), (attr, tagged, compress)
            assert g.path(0)[0] == "ref" and np.array_equal(g.path(0)[1], want.path(0)[1])
    raw = _vg_stream([chunk1, chunk2])
    for bad in (raw[:-3], raw[:40], _vg_stream([chunk1]).replace(b"\x02VG", b"\x03GAM", 1)):
        f = tmp_path / "bad.vg"
        f.write_bytes(bad)
        with pytest.raises(capi.PsiError):
            capi.Graph.load(f)
    with pytest.raises(capi.PsiError) as e:
        capi.Graph.load(tmp_path / "missing.vg")
    assert e.value.code == capi.ERR_IO
