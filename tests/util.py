"""Shared helpers of the test-suite: fixture loading and small synthetic graphs."""
from __future__ import annotations

import gzip
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
if os.fspath(ROOT) not in sys.path:
    sys.path.insert(0, os.fspath(ROOT))

from psi_b200 import capi  # noqa: E402


def golden_index() -> dict:
    with open(GOLDEN / "golden.json") as f:
        return json.load(f)


def read_fasta(path):
    """FASTA/FASTQ(.gz) -> (read_ptr u64[n+1], bases u8) through the library's reader."""
    r = capi.Reader(path)
    out = r.next(0)
    r.close()
    if out is None:
        return np.zeros(1, np.uint64), np.zeros(0, np.uint8)
    _, read_ptr, bases, _ = out
    return read_ptr, bases


def md5_tuples(t: np.ndarray) -> str:
    import hashlib
    return hashlib.md5(np.ascontiguousarray(t, dtype="<u8").tobytes()).hexdigest()


# ----------------------------------------------------------- synthetic --

class SplitMix:
    """Deterministic, version-independent generator (python's random module is
    avoided so that committed fixtures can be regenerated bit-for-bit)."""

    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def below(self, n):
        return self.next() % n

    def unit(self):
        return (self.next() >> 11) / float(1 << 53)


def random_bubble_gfa(seed, backbone=3000, sites=120, p_snp=0.8, p_ins=0.1, max_indel=6, n_frac=0.0,
                      multi_allele=0.1):
    """GFA1 text of a chain of bubbles (SNPs, insertions, deletions) with one P line,
    the graph family of the survey's fuzz loop."""
    rng = SplitMix(seed)
    acgt = "ACGT"
    seq = [acgt[rng.below(4)] for _ in range(backbone)]
    if n_frac > 0:
        for i in range(backbone):
            if rng.unit() < n_frac:
                seq[i] = "N"
    pos = sorted({1 + rng.below(backbone - 2) for _ in range(sites)})
    # keep sites apart by at least 1 backbone base so that bubbles do not nest
    kept, last = [], -2
    for p in pos:
        if p >= last + 2:
            kept.append(p)
            last = p
    segs, links, path = [], [], []
    nid = 0

    def new_seg(s):
        nonlocal nid
        nid += 1
        segs.append((nid, s))
        return nid

    prev_tails = []  # node ids whose out-edges go to the next segment
    cur = 0
    for p in kept:
        if p > cur:
            ref = new_seg("".join(seq[cur:p]))
            for t in prev_tails:
                links.append((t, ref))
            prev_tails = [ref]
            path.append(ref)
        u = rng.unit()
        if u < p_snp:  # SNP: ref base vs 1..2 alternative bases
            a = new_seg(seq[p])
            alts = [b for b in acgt if b != seq[p]]
            n_alt = 2 if rng.unit() < multi_allele else 1
            alt_ids = []
            for _ in range(n_alt):
                b = alts.pop(rng.below(len(alts)))
                alt_ids.append(new_seg(b))
            for t in prev_tails:
                links.append((t, a))
                for x in alt_ids:
                    links.append((t, x))
            prev_tails = [a] + alt_ids
            path.append(a)
            cur = p + 1
        elif u < p_snp + p_ins:  # insertion: optional extra node
            ins = new_seg("".join(acgt[rng.below(4)] for _ in range(1 + rng.below(max_indel))))
            for t in prev_tails:
                links.append((t, ins))
            prev_tails = prev_tails + [ins]
            cur = p
        else:  # deletion: the next few backbone bases are optional
            ln = 1 + rng.below(max_indel)
            ln = min(ln, backbone - p - 1)
            if ln <= 0:
                cur = p
                continue
            d = new_seg("".join(seq[p:p + ln]))
            for t in prev_tails:
                links.append((t, d))
            prev_tails = prev_tails + [d]
            path.append(d)
            cur = p + ln
    if cur < backbone:
        ref = new_seg("".join(seq[cur:]))
        for t in prev_tails:
            links.append((t, ref))
        path.append(ref)
    lines = ["H\tVN:Z:1.0"]
    lines += [f"S\t{i}\t{s}" for i, s in segs]
    # de-duplicate links, keep order
    seen = set()
    for a, b in links:
        if (a, b) not in seen:
            seen.add((a, b))
            lines.append(f"L\t{a}\t+\t{b}\t+\t0M")
    lines.append("P\tref\t" + ",".join(f"{i}+" for i in path) + "\t*")
    return "\n".join(lines) + "\n"


def random_walk_reads(g, n_reads, length, seed, n_frac=0.0):
    """Error-free reads: uniform start base, uniform out-edge at every node end
    (SURVEY 8d); walks that hit a sink early are discarded.  Returns (read_ptr, bases)."""
    rng = SplitMix(seed)
    out = []
    tries = 0
    lens = g.seq_start[1:] - g.seq_start[:-1]
    while len(out) < n_reads and tries < n_reads * 50:
        tries += 1
        pos = rng.below(int(g.n_bases))
        v = int(np.searchsorted(g.seq_start, pos, side="right") - 1)
        while lens[v] == 0:
            v += 1
        o = pos - int(g.seq_start[v])
        buf = bytearray()
        ok = True
        while len(buf) < length:
            s, e = int(g.seq_start[v]) + o, int(g.seq_start[v + 1])
            take = min(e - s, length - len(buf))
            buf += g.seq[s:s + take].tobytes()
            if len(buf) == length:
                break
            b, en = int(g.row_ptr[v]), int(g.row_ptr[v + 1])
            if b == en:
                ok = False
                break
            v = int(g.col[b + rng.below(en - b)])
            o = 0
        if ok:
            if n_frac > 0:
                for i in range(length):
                    if rng.unit() < n_frac:
                        buf[i] = ord("N")
            out.append(bytes(buf))
    read_ptr = np.zeros(len(out) + 1, np.uint64)
    read_ptr[1:] = np.cumsum([len(r) for r in out])
    bases = np.frombuffer(b"".join(out), np.uint8).copy() if out else np.zeros(0, np.uint8)
    return read_ptr, bases


def write_fasta(path, read_ptr, bases, gz=False):
    op = gzip.open if gz else open
    with op(path, "wt") as f:
        for i in range(len(read_ptr) - 1):
            f.write(f">r{i}\n{bases[int(read_ptr[i]):int(read_ptr[i + 1])].tobytes().decode()}\n")


def all_loci(g):
    lens = (g.seq_start[1:] - g.seq_start[:-1]).astype(np.int64)
    node = np.repeat(np.arange(g.n_nodes, dtype=np.uint32), lens)
    off = (np.arange(int(g.n_bases), dtype=np.int64) - np.repeat(g.seq_start[:-1].astype(np.int64), lens)).astype(np.uint32)
    return node, off
