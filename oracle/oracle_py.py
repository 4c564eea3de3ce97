"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product (psi_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "liboracle.so"
REF_DRIVER = _HERE / "_ref" / "psi_ref_driver"


class _Graph(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("seq_start", C.c_void_p), ("seq", C.c_void_p),
                ("row_ptr", C.c_void_p), ("col", C.c_void_p), ("node_id", C.c_void_p)]


class _Reads(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("read_ptr", C.c_void_p), ("bases", C.c_void_p),
                ("first_read_id", C.c_uint64)]


class _Paths(C.Structure):
    _fields_ = [("n_paths", C.c_uint64), ("path_ptr", C.c_void_p), ("nodes", C.c_void_p),
                ("head_off", C.c_void_p), ("tail_trim", C.c_void_p)]


class _Result(C.Structure):
    _fields_ = [("tuples", C.POINTER(C.c_uint64)), ("n", C.c_uint64), ("n_raw", C.c_uint64)]


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", os.fspath(_HERE), "oracle"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        L = C.CDLL(os.fspath(LIB_PATH))
        L.psi_oracle_seeding.restype = C.c_uint64
        L.psi_oracle_seeding.argtypes = [C.POINTER(_Reads), C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint64]
        L.psi_oracle_kmer_exact_matches.restype = C.c_uint64
        L.psi_oracle_kmer_exact_matches.argtypes = [C.POINTER(C.c_char_p), C.c_uint64, C.POINTER(C.c_char_p),
                                                    C.c_uint64, C.c_uint]
        L.psi_oracle_increment_kmer.restype = C.c_uint64
        L.psi_oracle_increment_kmer.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64]
        L.psi_oracle_seeds_on_paths.argtypes = [C.POINTER(_Graph), C.POINTER(_Paths), C.POINTER(_Reads), C.c_uint,
                                                C.c_uint, C.POINTER(_Result)]
        L.psi_oracle_seeds_off_paths.argtypes = [C.POINTER(_Graph), C.c_uint64, C.c_void_p, C.c_void_p,
                                                 C.POINTER(_Reads), C.c_uint, C.c_uint, C.POINTER(_Result)]
        L.psi_oracle_seeds_all.argtypes = [C.POINTER(_Graph), C.POINTER(_Paths), C.c_uint64, C.c_void_p, C.c_void_p,
                                           C.POINTER(_Reads), C.c_uint, C.c_uint, C.POINTER(_Result)]
        L.psi_oracle_seeds_closed_form.argtypes = [C.POINTER(_Graph), C.POINTER(_Reads), C.c_uint, C.c_uint,
                                                   C.POINTER(_Result)]
        L.psi_oracle_uncovered_loci.argtypes = [C.POINTER(_Graph), C.POINTER(_Paths), C.c_uint, C.c_uint,
                                                C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.psi_oracle_free.argtypes = [C.c_void_p]
        L.psi_oracle_free.restype = None
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OGraph:
    """Graph arrays for the oracle; `g` is any object with seq_start/seq/row_ptr/col and an id array."""

    def __init__(self, seq_start, seq, row_ptr, col, node_id):
        self.seq_start = np.ascontiguousarray(seq_start, np.uint64)
        self.seq = np.ascontiguousarray(seq, np.uint8)
        self.row_ptr = np.ascontiguousarray(row_ptr, np.uint64)
        self.col = np.ascontiguousarray(col, np.uint32)
        self.node_id = np.ascontiguousarray(node_id, np.uint64)
        self.c = _Graph(len(self.node_id), _p(self.seq_start), _p(self.seq), _p(self.row_ptr), _p(self.col),
                        _p(self.node_id))

    @classmethod
    def of(cls, g, ids="coord"):
        return cls(g.seq_start, g.seq, g.row_ptr, g.col, g.coord_id if ids == "coord" else g.internal_id)


class OReads:
    def __init__(self, read_ptr, bases, first_read_id=0):
        self.read_ptr = np.ascontiguousarray(read_ptr, np.uint64)
        self.bases = np.ascontiguousarray(bases, np.uint8)
        self.c = _Reads(len(self.read_ptr) - 1, _p(self.read_ptr), _p(self.bases), first_read_id)


class OPaths:
    def __init__(self, path_ptr, nodes, head_off=None, tail_trim=None):
        self.path_ptr = np.ascontiguousarray(path_ptr, np.uint64)
        self.nodes = np.ascontiguousarray(nodes, np.uint32)
        n = len(self.path_ptr) - 1
        self.head_off = np.zeros(n, np.uint32) if head_off is None else np.ascontiguousarray(head_off, np.uint32)
        self.tail_trim = np.zeros(n, np.uint32) if tail_trim is None else np.ascontiguousarray(tail_trim, np.uint32)
        self.c = _Paths(n, _p(self.path_ptr), _p(self.nodes), _p(self.head_off), _p(self.tail_trim))


def _take(res: _Result):
    n = res.n
    out = np.ctypeslib.as_array(res.tuples, shape=(n * 4,)).copy().reshape(n, 4) if n else np.zeros((0, 4), np.uint64)
    if res.tuples:
        lib().psi_oracle_free(res.tuples)
    return out, res.n_raw


def seeding(reads: OReads, k, d):
    n = lib().psi_oracle_seeding(C.byref(reads.c), k, d, None, None, 0)
    rid = np.zeros(n, np.uint64)
    off = np.zeros(n, np.uint64)
    lib().psi_oracle_seeding(C.byref(reads.c), k, d, _p(rid), _p(off), n)
    return rid, off


def kmer_exact_matches(set1, set2, k) -> int:
    a = (C.c_char_p * len(set1))(*[s.encode() for s in set1])
    b = (C.c_char_p * len(set2))(*[s.encode() for s in set2])
    return lib().psi_oracle_kmer_exact_matches(a, len(set1), b, len(set2), k)


def increment_kmer(kmer: str, pos: int):
    buf = C.create_string_buffer(kmer.encode())
    r = lib().psi_oracle_increment_kmer(buf, len(kmer), pos)
    return buf.value.decode(), r


def seeds_on_paths(g: OGraph, p: OPaths, r: OReads, k, d):
    res = _Result()
    rc = lib().psi_oracle_seeds_on_paths(C.byref(g.c), C.byref(p.c), C.byref(r.c), k, d, C.byref(res))
    assert rc == 0, rc
    return _take(res)


def seeds_off_paths(g: OGraph, loci_node, loci_off, r: OReads, k, d):
    ln = np.ascontiguousarray(loci_node, np.uint32)
    lo = np.ascontiguousarray(loci_off, np.uint32)
    res = _Result()
    rc = lib().psi_oracle_seeds_off_paths(C.byref(g.c), len(ln), _p(ln), _p(lo), C.byref(r.c), k, d, C.byref(res))
    assert rc == 0, rc
    return _take(res)


def seeds_all(g: OGraph, p: OPaths, loci_node, loci_off, r: OReads, k, d):
    ln = np.ascontiguousarray(loci_node, np.uint32)
    lo = np.ascontiguousarray(loci_off, np.uint32)
    res = _Result()
    rc = lib().psi_oracle_seeds_all(C.byref(g.c), C.byref(p.c) if p is not None else None, len(ln), _p(ln), _p(lo),
                                    C.byref(r.c), k, d, C.byref(res))
    assert rc == 0, rc
    return _take(res)


def seeds_closed_form(g: OGraph, r: OReads, k, d):
    res = _Result()
    rc = lib().psi_oracle_seeds_closed_form(C.byref(g.c), C.byref(r.c), k, d, C.byref(res))
    assert rc == 0, rc
    return _take(res)


def uncovered_loci(g: OGraph, p: OPaths, k, step=1):
    pn, po, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
    rc = lib().psi_oracle_uncovered_loci(C.byref(g.c), C.byref(p.c) if p is not None else None, k, step,
                                         C.byref(pn), C.byref(po), C.byref(n))
    assert rc == 0, rc
    cnt = n.value
    node = np.frombuffer(C.string_at(pn, cnt * 4), np.uint32).copy() if cnt else np.zeros(0, np.uint32)
    off = np.frombuffer(C.string_at(po, cnt * 4), np.uint32).copy() if cnt else np.zeros(0, np.uint32)
    lib().psi_oracle_free(pn)
    lib().psi_oracle_free(po)
    return node, off


# ---------------------------------------------------------------- reference --

def have_reference() -> bool:
    return REF_DRIVER.exists() and os.access(REF_DRIVER, os.X_OK)


def run_reference(gfa, fastq, k, d=0, n_paths=16, patched=True, context=0, chunk=0, first_read=0, max_reads=0,
                  want_out=True, extra=(), timeout=3600):
    """Runs the compiled, unmodified reference (oracle/_ref/psi_ref_driver).  Returns (tuples or None, stats)."""
    assert have_reference(), "oracle/_ref/psi_ref_driver missing: `make -C oracle ref` where /root/reference exists"
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "out.bin")
        cmd = [os.fspath(REF_DRIVER), "--gfa", os.fspath(gfa), "-k", str(k), "-d", str(d), "-n", str(n_paths),
               "-t", str(context), "-c", str(chunk)]
        if fastq:
            cmd += ["--fastq", os.fspath(fastq)]
        if not patched:
            cmd.append("-P")
        if first_read:
            cmd += ["--first-read", str(first_read)]
        if max_reads:
            cmd += ["--max-reads", str(max_reads)]
        if want_out:
            cmd += ["--out", out]
        cmd += list(extra)
        env = dict(os.environ, TMPDIR=td, OMP_PROC_BIND="false", OMP_NUM_THREADS="1")
        pr = subprocess.run(cmd, check=True, capture_output=True, text=True, env=env, timeout=timeout)
        stats = json.loads(pr.stdout.strip().splitlines()[-1])
        tuples = np.fromfile(out, np.uint64).reshape(-1, 4) if want_out else None
        return tuples, stats


# ---------------------------------------------------------------- MEM mode --

def path_texts(g, path_ptr, nodes, head, tail):
    """Forward text of every indexed path and, per text position, the global graph position it comes from
    (sequence(path, Forward), reference include/psi/path_interface.hpp:207-255)."""
    texts, gpos = [], []
    for p in range(len(head)):
        ns = nodes[int(path_ptr[p]):int(path_ptr[p + 1])]
        seq = np.concatenate([g.seq[int(g.seq_start[v]):int(g.seq_start[v + 1])] for v in ns]) if len(ns) else np.zeros(0, np.uint8)
        gp = np.concatenate([np.arange(int(g.seq_start[v]), int(g.seq_start[v + 1]), dtype=np.int64) for v in ns]) if len(ns) else np.zeros(0, np.int64)
        lo, hi = int(head[p]), len(seq) - int(tail[p])
        texts.append(seq[lo:hi].tobytes().upper())
        gpos.append(gp[lo:hi])
    return texts, gpos


def _occurrences(texts, pat):
    out = []
    for ti, t in enumerate(texts):
        i = t.find(pat)
        while i != -1:
            out.append((ti, i))
            i = t.find(pat, i + 1)
    return out


def find_mems(texts, pattern: bytes, minlen: int, gocc_threshold=0, max_mem=0):
    """Restatement of find_mems (reference include/psi/index_iter.hpp:854-906) as called by
    SeedFinder::seeds_on_paths(sequence, callback) (seed_finder.hpp:1459-1479): a greedy left-to-right scan that
    extends pattern[start : start + plen] while it still occurs in the path text, reports ALL its occurrences as soon as
    it is at least minlen long and occurs at most gocc_threshold times, then restarts one character further on; a
    character that cannot be appended (or an 'N') also restarts the scan behind it.  Returns the raw hits in emission
    order: (start, plen, gocc, text index, text offset)."""
    if gocc_threshold == 0:
        gocc_threshold = 2 ** 32 - 1
    if max_mem == 0:
        max_mem = 2 ** 32 - 1
    pattern = pattern.upper()
    hits, start, plen, has_hit, nof = [], 0, 0, False, 0
    occ = None                                    # occurrences of pattern[start:start+plen]; None = all positions (plen 0)
    while start + plen < len(pattern):
        if plen >= minlen and len(occ) <= gocc_threshold:
            has_hit = True
            for ti, o in occ:
                hits.append((start, plen, len(occ), ti, o))
                nof += 1
            if nof >= max_mem:
                break
        c = pattern[start + plen:start + plen + 1]
        nxt = None
        if not has_hit and c != b"N":
            nxt = _occurrences(texts, pattern[start:start + plen + 1])
        if has_hit or c == b"N" or not nxt:
            start, plen, has_hit, occ = start + plen + 1, 0, False, None
            continue
        occ = nxt
        plen += 1
    return hits


# ---------------------------------------------------- distance verification --

def verify_distance(g, v: int, o: int, u: int, p: int, dmin: int, dmax: int) -> bool:
    """Restatement of SeedFinder::verify_distance over DiVerG's distance index (reference seed_finder.hpp:1300-1317;
    index built by create_distance_index :1193-1265 = diverg::util::create_distance_index, dindex.hpp:767-914:
    D = A^dmin (A + I)^(dmax - dmin) over the character-level adjacency A, i.e. D[x][y] is set iff some walk of
    dmin <= l <= dmax character steps leads from character x to character y): is there such a walk from locus (v, o)
    to locus (u, p)?  v, u are 0-based node ranks.  Inside one node only the offsets count (:1306-1309).  Across nodes:
    breadth-first over (node, distance at which its first character is reached) states, bounded by dmax -- any graph,
    cycles included."""
    if v == u:
        return o <= p and dmin <= p - o <= dmax
    def length(x):
        return int(g.seq_start[x + 1] - g.seq_start[x])
    seen = set()
    todo = []
    def push(x, s):
        if s <= dmax and (x, s) not in seen:
            seen.add((x, s))
            todo.append((x, s))
    for e in range(int(g.row_ptr[v]), int(g.row_ptr[v + 1])):
        push(int(g.col[e]), length(v) - o)
    while todo:
        x, s = todo.pop()
        if x == u and dmin <= s + p <= dmax:
            return True
        for e in range(int(g.row_ptr[x]), int(g.row_ptr[x + 1])):
            push(int(g.col[e]), s + length(x))
    return False
