"""ctypes binding of bench_support/libpsi_synth.so (synthetic graphs/reads)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = _HERE / "libpsi_synth.so"
        if not so.exists():
            subprocess.run(["make", "-s", "-C", os.fspath(_HERE)], check=True)
        L = C.CDLL(os.fspath(so))
        L.psi_synth_graph_create.restype = C.c_void_p
        L.psi_synth_graph_create.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_uint64,
                                             C.c_uint64]
        L.psi_synth_graph_sizes.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 4
        L.psi_synth_graph_sizes.restype = None
        L.psi_synth_graph_fill.argtypes = [C.c_void_p] * 7
        L.psi_synth_graph_fill.restype = None
        L.psi_synth_graph_free.argtypes = [C.c_void_p]
        L.psi_synth_graph_free.restype = None
        L.psi_synth_reads.restype = C.c_uint64
        L.psi_synth_reads.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                      C.c_uint32, C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def whole_genome_components(total_bp, total_sites, n_components=24):
    """BASELINE configs[4]'s recipe: n components of equal size, the chr22 site mix, seeds 100 + c."""
    return [dict(backbone=total_bp // n_components, sites=total_sites // n_components, p_snp=0.90, p_ins=0.05, tri_frac=0.0,
                 seeds=(100 + c, 200 + c)) for c in range(n_components)]


# named shapes of BASELINE.json / SURVEY 8d
SHAPES = {
    # backbone, sites, p_snp, p_ins, tri_frac, seeds
    "chr22": dict(backbone=51_000_000, sites=1_000_000, p_snp=0.90, p_ins=0.05, tri_frac=0.0, seeds=(22, 23)),
    "chr22_1_51": dict(backbone=1_000_000, sites=19_608, p_snp=0.90, p_ins=0.05, tri_frac=0.0, seeds=(22, 23)),
    # a chromosome-1-sized instance of the chr22 recipe (5x): exercises the grouped index build (16 paths x 255 Mbp = 4 G
    # path windows) and a 5.4 GB index
    "chr1": dict(backbone=250_000_000, sites=4_900_000, p_snp=0.90, p_ins=0.05, tri_frac=0.0, seeds=(1, 2)),
    # BASELINE configs[4] as written: 3.1 Gbp, 80 M sites, 24 components (a sliced index build: more than 2^32 pairs)
    "wg": dict(components=whole_genome_components(3_100_000_000, 80_000_000)),
    # ... and at 1/4 and 1/16 of its size
    "wg_1_4": dict(components=whole_genome_components(775_000_000, 20_000_000)),
    "wg_1_16": dict(components=whole_genome_components(193_750_000, 5_000_000)),
    "mhc": dict(backbone=5_000_000, sites=416_667, p_snp=1.0, p_ins=0.0, tri_frac=0.10, seeds=(6, 7)),
}


def graph_arrays(backbone, sites, p_snp=0.9, p_ins=0.05, tri_frac=0.0, seeds=(22, 23)):
    """Returns dict(ids, seq_start, seq, row_ptr, col, path_ptr, path_nodes) in generation (backbone) order."""
    L = lib()
    h = L.psi_synth_graph_create(backbone, sites, p_snp, p_ins, tri_frac, seeds[0], seeds[1])
    n, m, b, pl = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    L.psi_synth_graph_sizes(h, C.byref(n), C.byref(m), C.byref(b), C.byref(pl))
    ids = np.zeros(n.value, np.uint64)
    seq_start = np.zeros(n.value + 1, np.uint64)
    seq = np.zeros(b.value, np.uint8)
    row_ptr = np.zeros(n.value + 1, np.uint64)
    col = np.zeros(max(m.value, 1), np.uint32)
    path = np.zeros(max(pl.value, 1), np.uint32)
    L.psi_synth_graph_fill(h, _p(ids), _p(seq_start), _p(seq), _p(row_ptr), _p(col), _p(path))
    L.psi_synth_graph_free(h)
    return dict(ids=ids, seq_start=seq_start, seq=seq, row_ptr=row_ptr, col=col[: m.value],
                path_ptr=np.array([0, pl.value], np.uint64), path_nodes=path[: pl.value])


def graph_arrays_multi(components):
    """Several independent components (one embedded path each) as ONE graph: `components` is a list of graph_arrays
    keyword dicts.  Node ids continue from component to component; there is no edge between components."""
    parts = [graph_arrays(**c) for c in components]
    ids, seq_start, seq, row_ptr, col, path_ptr, path_nodes = [], [np.zeros(1, np.uint64)], [], [np.zeros(1, np.uint64)], [], [0], []
    n0 = b0 = m0 = 0
    for a in parts:
        n, b, m = len(a["ids"]), len(a["seq"]), len(a["col"])
        ids.append(a["ids"] + np.uint64(n0))
        seq_start.append(a["seq_start"][1:] + np.uint64(b0))
        seq.append(a["seq"])
        row_ptr.append(a["row_ptr"][1:] + np.uint64(m0))
        col.append(a["col"] + np.uint32(n0))
        path_nodes.append(a["path_nodes"] + np.uint32(n0))
        path_ptr.append(path_ptr[-1] + len(a["path_nodes"]))
        n0, b0, m0 = n0 + n, b0 + b, m0 + m
    return dict(ids=np.concatenate(ids), seq_start=np.concatenate(seq_start), seq=np.concatenate(seq), row_ptr=np.concatenate(row_ptr),
                col=np.concatenate(col), path_ptr=np.array(path_ptr, np.uint64), path_nodes=np.concatenate(path_nodes))


def reads(g, n_reads, length, seed):
    """g: any object with n_nodes, seq_start, seq, row_ptr, col.  Returns (read_ptr u64, bases u8)."""
    bases = np.zeros(n_reads * length, np.uint8)
    seq_start = np.ascontiguousarray(g.seq_start, np.uint64)
    seq = np.ascontiguousarray(g.seq, np.uint8)
    row_ptr = np.ascontiguousarray(g.row_ptr, np.uint64)
    col = np.ascontiguousarray(g.col, np.uint32)
    done = lib().psi_synth_reads(len(seq_start) - 1, _p(seq_start), _p(seq), _p(row_ptr), _p(col), n_reads, length,
                                 seed, _p(bases))
    assert done == n_reads, f"only {done} of {n_reads} reads could be drawn"
    read_ptr = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(length)
    return read_ptr, bases
