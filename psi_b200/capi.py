"""ctypes binding of libpsi_b200.so (include/psi_b200.h).

This is plumbing for tests/ and bench.py: numpy arrays in, numpy arrays out.
The product is the shared library; nothing here computes anything, and there is
no CPU fallback -- if the library (or a GPU, for the device entry points) is
missing, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libpsi_b200.so"

OK = 0
ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_IO, ERR_STATE, ERR_OVERFLOW = -1, -2, -3, -4, -5, -6
ON_PATHS, OFF_PATHS, ALL, SORTED, NO_RESOLVE, COMPACT, DENSE, DENSE5 = 1, 2, 3, 4, 8, 16, 32, 64
NIL32 = 0xFFFFFFFF

# every symbol include/psi_b200.h declares
SYMBOLS = [
    "psi_b200_graph_load_gfa", "psi_b200_graph_load_vg", "psi_b200_graph_load", "psi_b200_graph_from_arrays", "psi_b200_graph_free",
    "psi_b200_graph_get_view", "psi_b200_graph_path", "psi_b200_graph_write_gfa",
    "psi_b200_pick_paths", "psi_b200_pathset_free", "psi_b200_pathset_get_view", "psi_b200_pathset_load_reference",
    "psi_b200_reader_open", "psi_b200_reader_next", "psi_b200_reader_close",
    "psi_b200_reader_next_packed", "psi_b200_pack_bases", "psi_b200_submit_chunk_packed",
    "psi_b200_seeds_all_async", "psi_b200_wait", "psi_b200_fetch_dense", "psi_b200_fetch_dense_async", "psi_b200_dense_counts", "psi_b200_dense_layout", "psi_b200_dense5_layout",
    "psi_b200_build_mem_index", "psi_b200_find_mems", "psi_b200_fetch_mems",
    "psi_b200_create_distance_index", "psi_b200_verify_distance",
    "psi_b200_global_error",
    "psi_b200_create", "psi_b200_fork", "psi_b200_destroy", "psi_b200_last_error", "psi_b200_set_stream", "psi_b200_sync", "psi_b200_set_option",
    "psi_b200_set_graph", "psi_b200_set_paths", "psi_b200_find_loci", "psi_b200_get_loci", "psi_b200_set_loci",
    "psi_b200_submit_chunk", "psi_b200_submit_chunk_device", "psi_b200_seeds_all",
    "psi_b200_fetch", "psi_b200_fetch32", "psi_b200_fetch_kinds", "psi_b200_fetch_device", "psi_b200_host_alloc", "psi_b200_host_free",
    "psi_b200_counters", "psi_b200_reset_counters", "psi_b200_version",
]


class PsiError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"psi_b200 error {code}: {msg}")
        self.code = code


class GraphView(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("n_edges", C.c_uint64), ("n_bases", C.c_uint64), ("n_paths", C.c_uint64),
                ("seq_start", C.POINTER(C.c_uint64)), ("seq", C.POINTER(C.c_char)),
                ("row_ptr", C.POINTER(C.c_uint64)), ("col", C.POINTER(C.c_uint32)),
                ("internal_id", C.POINTER(C.c_uint64)), ("coord_id", C.POINTER(C.c_uint64))]


class PathSetView(C.Structure):
    _fields_ = [("n_paths", C.c_uint64), ("path_ptr", C.POINTER(C.c_uint64)), ("nodes", C.POINTER(C.c_uint32)),
                ("head_off", C.POINTER(C.c_uint32)), ("tail_trim", C.POINTER(C.c_uint32))]


class ChunkView(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("first_read_id", C.c_uint64),
                ("read_ptr", C.POINTER(C.c_uint64)), ("bases", C.POINTER(C.c_char)),
                ("name_ptr", C.POINTER(C.c_uint64)), ("names", C.POINTER(C.c_char))]


class PackedChunk(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("first_read_id", C.c_uint64), ("n_bases", C.c_uint64),
                ("read_len", C.c_uint32), ("reserved", C.c_uint32),
                ("read_ptr", C.c_void_p), ("words", C.c_void_p), ("exc", C.c_void_p), ("n_exc", C.c_uint64),
                ("name_ptr", C.c_void_p), ("names", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("n_edges", C.c_uint64), ("n_bases", C.c_uint64),
                ("n_path_bases", C.c_uint64), ("n_index_entries", C.c_uint64), ("n_index_kmers", C.c_uint64),
                ("index_bytes", C.c_uint64), ("index_buckets", C.c_uint64),
                ("index_slot_bytes", C.c_uint32), ("index_stash_used", C.c_uint32),
                ("n_offpath_entries", C.c_uint64), ("n_offpath_walks", C.c_uint64),
                ("offpath_mode", C.c_uint32), ("fused", C.c_uint32),
                ("n_loci", C.c_uint64), ("n_reads", C.c_uint64), ("n_seeds", C.c_uint64),
                ("n_hits_on", C.c_uint64), ("n_hits_off", C.c_uint64), ("n_hits", C.c_uint64),
                ("n_walks", C.c_uint64), ("n_on_probe_sectors", C.c_uint64),
                ("ms_index_build", C.c_float), ("ms_find_loci", C.c_float),
                ("ms_h2d", C.c_float), ("ms_pack", C.c_float), ("ms_read_index", C.c_float), ("ms_on", C.c_float),
                ("ms_off", C.c_float), ("ms_resolve", C.c_float), ("ms_sort", C.c_float), ("ms_d2h", C.c_float),
                ("launches", C.c_uint32), ("ms_probe", C.c_float),
                ("ms_probe_sum", C.c_double), ("ms_on_sum", C.c_double), ("timed_steps", C.c_uint64),
                ("n_gocc_dropped", C.c_uint64), ("code_by_rank", C.c_uint32), ("code_off_bits", C.c_uint32),
                ("n_dindex_entries", C.c_uint64), ("dindex_bytes", C.c_uint64), ("dindex_mode", C.c_uint32), ("ms_dindex_build", C.c_float),
                ("index_build_slices", C.c_uint32), ("reserved0", C.c_uint32)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_ if not n.startswith("reserved")}


_lib = None
HOST_LIB_PATH = _HERE / "libpsi_b200_host.so"
# the host half of the C-ABI (graphs, paths, reads): what libpsi_b200_host.so exports
HOST_SYMBOLS = [s_ for s_ in SYMBOLS if s_.startswith(("psi_b200_graph_", "psi_b200_pick_paths", "psi_b200_pathset_", "psi_b200_reader_",
                                                       "psi_b200_pack_bases", "psi_b200_global_error"))]
_host_only = False


def use_host_library():
    """Bind libpsi_b200_host.so instead of libpsi_b200.so: graphs, paths and reads only, no CUDA runtime mapped into the
    process (bench.py's reference arm builds its inputs with it).  Must be called before the first lib()."""
    global _host_only
    assert _lib is None, "use_host_library() must come before the first call into the library"
    _host_only = True


def lib() -> C.CDLL:
    """Load libpsi_b200.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = HOST_LIB_PATH if _host_only else LIB_PATH
    if not path.exists():
        raise PsiError(ERR_IO, f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
    L = C.CDLL(os.fspath(path))
    if _host_only:
        _bind_host(L)
        _lib = L
        return L
    u64p, u32p, vp = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_void_p
    _bind_host(L)
    L.psi_b200_version.restype = C.c_char_p
    L.psi_b200_last_error.restype = C.c_char_p
    L.psi_b200_last_error.argtypes = [vp]
    _bind_device(L, u64p, u32p, vp)
    _lib = L
    return L


def _bind_host(L):
    u64p, u32p, vp = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_void_p
    L.psi_b200_global_error.restype = C.c_char_p
    L.psi_b200_graph_load_gfa.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.psi_b200_graph_load_vg.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.psi_b200_graph_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.psi_b200_graph_from_arrays.argtypes = [C.c_uint64, vp, vp, vp, vp, vp, C.c_uint64, vp, vp, C.c_int, C.POINTER(vp)]
    L.psi_b200_graph_free.argtypes = [vp]
    L.psi_b200_graph_free.restype = None
    L.psi_b200_graph_get_view.argtypes = [vp, C.POINTER(GraphView)]
    L.psi_b200_graph_path.argtypes = [vp, C.c_uint64, C.POINTER(C.c_char_p), C.POINTER(u32p), u64p]
    L.psi_b200_graph_write_gfa.argtypes = [vp, C.c_char_p]
    L.psi_b200_pick_paths.argtypes = [vp, C.c_uint, C.c_int, C.c_uint, C.c_uint64, C.POINTER(vp)]
    L.psi_b200_pathset_load_reference.argtypes = [vp, C.c_char_p, C.POINTER(vp), u64p]
    L.psi_b200_pathset_free.argtypes = [vp]
    L.psi_b200_pathset_free.restype = None
    L.psi_b200_pathset_get_view.argtypes = [vp, C.POINTER(PathSetView)]
    L.psi_b200_reader_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.psi_b200_reader_next.argtypes = [vp, C.c_uint64, C.POINTER(ChunkView)]
    L.psi_b200_reader_close.argtypes = [vp]
    L.psi_b200_reader_close.restype = None
    L.psi_b200_reader_next_packed.argtypes = [vp, C.c_uint64, C.POINTER(PackedChunk)]
    L.psi_b200_pack_bases.argtypes = [vp, C.c_uint64, vp, vp, C.c_uint64, u64p]


def _bind_device(L, u64p, u32p, vp):
    L.psi_b200_submit_chunk_packed.argtypes = [vp, C.POINTER(PackedChunk), C.c_uint, C.c_int]
    L.psi_b200_seeds_all_async.argtypes = [vp, C.c_uint]
    L.psi_b200_wait.argtypes = [vp, u64p]
    L.psi_b200_fetch_dense.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, u64p, u64p]
    L.psi_b200_fetch_dense_async.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64]
    L.psi_b200_dense_counts.argtypes = [vp, u64p, u64p]
    L.psi_b200_dense_layout.argtypes = [vp, C.POINTER(C.c_uint)]
    L.psi_b200_dense5_layout.argtypes = [vp, C.POINTER(C.c_uint), C.POINTER(C.c_int)]
    L.psi_b200_build_mem_index.argtypes = [vp, C.c_uint64, vp, vp, vp, vp]
    L.psi_b200_find_mems.argtypes = [vp, C.c_uint, u64p]
    L.psi_b200_fetch_mems.argtypes = [vp, vp, C.c_uint64, u64p]
    L.psi_b200_create_distance_index.argtypes = [vp, C.c_uint, C.c_uint]
    L.psi_b200_verify_distance.argtypes = [vp, C.c_uint64, vp, vp, C.c_int]
    L.psi_b200_create.argtypes = [C.c_int, C.c_uint, C.POINTER(vp)]
    L.psi_b200_fork.argtypes = [vp, C.POINTER(vp)]
    L.psi_b200_destroy.argtypes = [vp]
    L.psi_b200_destroy.restype = None
    L.psi_b200_set_stream.argtypes = [vp, vp]
    L.psi_b200_sync.argtypes = [vp]
    L.psi_b200_set_option.argtypes = [vp, C.c_char_p, C.c_longlong]
    L.psi_b200_set_graph.argtypes = [vp, C.c_uint64, vp, vp, vp, vp, vp]
    L.psi_b200_set_paths.argtypes = [vp, C.c_uint64, vp, vp, vp, vp]
    L.psi_b200_find_loci.argtypes = [vp, C.c_uint, u64p]
    L.psi_b200_get_loci.argtypes = [vp, vp, vp, C.c_uint64, u64p]
    L.psi_b200_set_loci.argtypes = [vp, C.c_uint64, vp, vp]
    L.psi_b200_submit_chunk.argtypes = [vp, C.c_uint64, vp, vp, C.c_uint64, C.c_uint]
    L.psi_b200_submit_chunk_device.argtypes = [vp, C.c_uint64, vp, vp, C.c_uint64, C.c_uint64, C.c_uint]
    L.psi_b200_seeds_all.argtypes = [vp, C.c_uint, u64p]
    L.psi_b200_fetch.argtypes = [vp, vp, C.c_uint64, u64p]
    L.psi_b200_fetch32.argtypes = [vp, vp, C.c_uint64, u64p]
    L.psi_b200_fetch_kinds.argtypes = [vp, vp, C.c_uint64, u64p]
    L.psi_b200_fetch_device.argtypes = [vp, C.POINTER(vp), u64p]
    L.psi_b200_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.psi_b200_host_free.argtypes = [vp]
    L.psi_b200_host_free.restype = None
    L.psi_b200_counters.argtypes = [vp, C.POINTER(Counters)]
    L.psi_b200_reset_counters.argtypes = [vp]


def _check(rc: int, ctx=None):
    if rc != OK:
        L = lib()
        msg = L.psi_b200_last_error(ctx) if ctx else L.psi_b200_global_error()
        raise PsiError(rc, (msg or b"").decode())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _bytes(ptr, n):
    """n bytes at a char pointer as a numpy copy (any size: ctypes.string_at stops at 2 GiB)."""
    if n == 0:
        return np.zeros(0, np.uint8)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(int(n),)).copy()


def _np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype) if dtype != np.uint8 else _bytes(ptr, n)


class Graph:
    """Flattened sequence graph (host).  Arrays are numpy copies of the library's."""

    def __init__(self, handle):
        self._h = handle
        v = GraphView()
        _check(lib().psi_b200_graph_get_view(handle, C.byref(v)))
        n = v.n_nodes
        self.n_nodes, self.n_edges, self.n_bases, self.n_paths = n, v.n_edges, v.n_bases, v.n_paths
        self.seq_start = np.ctypeslib.as_array(v.seq_start, shape=(n + 1,)).copy()
        self.row_ptr = np.ctypeslib.as_array(v.row_ptr, shape=(n + 1,)).copy()
        self.col = np.ctypeslib.as_array(v.col, shape=(max(v.n_edges, 1),))[: v.n_edges].copy() \
            if v.n_edges else np.zeros(0, np.uint32)
        self.seq = _bytes(v.seq, v.n_bases)
        self.internal_id = np.ctypeslib.as_array(v.internal_id, shape=(n,)).copy()
        self.coord_id = np.ctypeslib.as_array(v.coord_id, shape=(n,)).copy()

    @classmethod
    def load_gfa(cls, path, sort=True) -> "Graph":
        h = C.c_void_p()
        _check(lib().psi_b200_graph_load_gfa(os.fspath(path).encode(), int(sort), C.byref(h)))
        return cls(h)

    @classmethod
    def load(cls, path, sort=True) -> "Graph":
        """By file name like gum::util::load: *.vg is read as vg (protobuf), anything else as GFA."""
        h = C.c_void_p()
        _check(lib().psi_b200_graph_load(os.fspath(path).encode(), int(sort), C.byref(h)))
        return cls(h)

    @classmethod
    def load_vg(cls, path, sort=True) -> "Graph":
        h = C.c_void_p()
        _check(lib().psi_b200_graph_load_vg(os.fspath(path).encode(), int(sort), C.byref(h)))
        return cls(h)

    @classmethod
    def from_arrays(cls, ids, seq_start, seq, row_ptr, col, path_ptr=None, path_nodes=None, sort=False) -> "Graph":
        ids = np.ascontiguousarray(ids, np.uint64)
        seq_start = np.ascontiguousarray(seq_start, np.uint64)
        seq = np.ascontiguousarray(seq, np.uint8)
        row_ptr = np.ascontiguousarray(row_ptr, np.uint64)
        col = np.ascontiguousarray(col, np.uint32)
        n_paths = 0
        if path_ptr is not None:
            path_ptr = np.ascontiguousarray(path_ptr, np.uint64)
            path_nodes = np.ascontiguousarray(path_nodes, np.uint32)
            n_paths = len(path_ptr) - 1
        h = C.c_void_p()
        _check(lib().psi_b200_graph_from_arrays(len(ids), _ptr(ids), _ptr(seq_start), _ptr(seq), _ptr(row_ptr),
                                                _ptr(col), n_paths, _ptr(path_ptr), _ptr(path_nodes), int(sort),
                                                C.byref(h)))
        return cls(h)

    def path(self, i):
        name, nodes, n = C.c_char_p(), C.POINTER(C.c_uint32)(), C.c_uint64()
        _check(lib().psi_b200_graph_path(self._h, i, C.byref(name), C.byref(nodes), C.byref(n)))
        return name.value.decode(), np.ctypeslib.as_array(nodes, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint32)

    def write_gfa(self, path):
        _check(lib().psi_b200_graph_write_gfa(self._h, os.fspath(path).encode()))

    def load_reference_paths(self, paths_file):
        """(PathSet, context) from a `<prefix>_paths` file saved by the reference."""
        h, ctx = C.c_void_p(), C.c_uint64()
        _check(lib().psi_b200_pathset_load_reference(self._h, os.fspath(paths_file).encode(), C.byref(h), C.byref(ctx)))
        return PathSet(h), ctx.value

    def pick_paths(self, n, patched=True, context=0, seed=1) -> "PathSet":
        h = C.c_void_p()
        _check(lib().psi_b200_pick_paths(self._h, n, int(patched), context, seed, C.byref(h)))
        return PathSet(h)

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:     # (module globals are gone at interpreter shutdown)
            lib().psi_b200_graph_free(self._h)
            self._h = None


class PathSet:
    def __init__(self, handle=None, path_ptr=None, nodes=None, head_off=None, tail_trim=None):
        self._h = handle
        if handle is not None:
            v = PathSetView()
            _check(lib().psi_b200_pathset_get_view(handle, C.byref(v)))
            n = v.n_paths
            self.path_ptr = np.ctypeslib.as_array(v.path_ptr, shape=(n + 1,)).copy()
            ne = int(self.path_ptr[-1])
            self.nodes = np.ctypeslib.as_array(v.nodes, shape=(ne,)).copy() if ne else np.zeros(0, np.uint32)
            self.head_off = np.ctypeslib.as_array(v.head_off, shape=(n,)).copy() if n else np.zeros(0, np.uint32)
            self.tail_trim = np.ctypeslib.as_array(v.tail_trim, shape=(n,)).copy() if n else np.zeros(0, np.uint32)
        else:
            self.path_ptr = np.ascontiguousarray(path_ptr, np.uint64)
            self.nodes = np.ascontiguousarray(nodes, np.uint32)
            n = len(self.path_ptr) - 1
            self.head_off = np.zeros(n, np.uint32) if head_off is None else np.ascontiguousarray(head_off, np.uint32)
            self.tail_trim = np.zeros(n, np.uint32) if tail_trim is None else np.ascontiguousarray(tail_trim, np.uint32)

    @property
    def n_paths(self):
        return len(self.path_ptr) - 1

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:     # (module globals are gone at interpreter shutdown)
            lib().psi_b200_pathset_free(self._h)
            self._h = None


class Reader:
    """FASTQ/FASTA chunk reader (readRecords, reference sequence.hpp:1608-1624)."""

    def __init__(self, path):
        self._h = C.c_void_p()
        _check(lib().psi_b200_reader_open(os.fspath(path).encode(), C.byref(self._h)))

    def next(self, max_reads=0):
        """Returns (first_read_id, read_ptr[u64 n+1], bases[u8], names) or None at end of input."""
        v = ChunkView()
        _check(lib().psi_b200_reader_next(self._h, max_reads, C.byref(v)))
        if v.n_reads == 0:
            return None
        n = v.n_reads
        read_ptr = np.ctypeslib.as_array(v.read_ptr, shape=(n + 1,)).copy()
        bases = _bytes(v.bases, int(read_ptr[-1]))
        name_ptr = np.ctypeslib.as_array(v.name_ptr, shape=(n + 1,)).copy()
        names_raw = C.string_at(v.names, int(name_ptr[-1]))
        names = [names_raw[name_ptr[i]:name_ptr[i + 1]].decode() for i in range(n)]
        return v.first_read_id, read_ptr, bases, names

    def next_packed(self, max_reads=0):
        """Returns a Packed chunk (2-bit words, numpy copies) or None at end of input."""
        v = PackedChunk()
        _check(lib().psi_b200_reader_next_packed(self._h, max_reads, C.byref(v)))
        if v.n_reads == 0:
            return None
        n = v.n_reads
        read_ptr = np.ctypeslib.as_array(C.cast(v.read_ptr, C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
        words = np.ctypeslib.as_array(C.cast(v.words, C.POINTER(C.c_uint64)), shape=(v.n_bases // 32 + 2,)).copy()
        exc = np.ctypeslib.as_array(C.cast(v.exc, C.POINTER(C.c_uint64)), shape=(v.n_exc,)).copy() if v.n_exc else np.zeros(0, np.uint64)
        return Packed(read_ptr, words, exc, v.read_len, v.first_read_id)

    def close(self):
        if self._h:
            lib().psi_b200_reader_close(self._h)
            self._h = None

    def __del__(self):
        if lib is not None:
            self.close()


class Packed:
    """A read chunk as 2-bit words (psi_b200_packed_chunk), host side."""

    def __init__(self, read_ptr, words, exc, read_len=0, first_read_id=0):
        self.read_ptr = np.ascontiguousarray(read_ptr, np.uint64)
        self.words = np.ascontiguousarray(words, np.uint64)
        self.exc = np.ascontiguousarray(exc, np.uint64)
        self.read_len = int(read_len)
        self.first_read_id = int(first_read_id)

    @property
    def n_reads(self):
        return len(self.read_ptr) - 1

    @property
    def n_bases(self):
        return int(self.read_ptr[-1])

    @classmethod
    def pack(cls, read_ptr, bases, first_read_id=0) -> "Packed":
        """psi_b200_pack_bases over an ASCII chunk."""
        read_ptr = np.ascontiguousarray(read_ptr, np.uint64)
        bases = np.ascontiguousarray(bases, np.uint8)
        n = int(read_ptr[-1])
        words = np.zeros(n // 32 + 2, np.uint64)
        n_exc = C.c_uint64()
        _check(lib().psi_b200_pack_bases(_ptr(bases), n, _ptr(words), None, 0, C.byref(n_exc)))
        exc = np.zeros(n_exc.value, np.uint64)
        if n_exc.value:
            _check(lib().psi_b200_pack_bases(_ptr(bases), n, _ptr(words), _ptr(exc), n_exc.value, C.byref(n_exc)))
        lens = np.diff(read_ptr.astype(np.int64))
        uniform = int(lens[0]) if len(lens) and (lens == lens[0]).all() and lens[0] > 0 else 0
        return cls(read_ptr, words, exc, uniform, first_read_id)

    def struct(self, with_read_ptr=None) -> PackedChunk:
        """with_read_ptr: None = omit the offsets when all reads have one length."""
        v = PackedChunk()
        v.n_reads, v.first_read_id, v.n_bases = self.n_reads, self.first_read_id, self.n_bases
        use_ptr = (self.read_len == 0) if with_read_ptr is None else with_read_ptr
        v.read_len = self.read_len
        v.read_ptr = self.read_ptr.ctypes.data if use_ptr else None
        v.words = self.words.ctypes.data
        v.exc = self.exc.ctypes.data if len(self.exc) else None
        v.n_exc = len(self.exc)
        return v


class Context:
    """One GPU context (psi_b200_ctx)."""

    def __init__(self, seed_len: int, device: int = 0, _handle=None):
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
        else:
            _check(lib().psi_b200_create(device, seed_len, C.byref(self._h)))
        self.k = seed_len

    def fork(self) -> "Context":
        """A second pipeline sharing this context's resident graph / index / loci."""
        h = C.c_void_p()
        self._ck(lib().psi_b200_fork(self._h, C.byref(h)))
        return Context(self.k, _handle=h)

    def _ck(self, rc):
        _check(rc, self._h)

    def close(self):
        if getattr(self, "_h", None):
            lib().psi_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        if lib is not None:     # (module globals are gone at interpreter shutdown)
            self.close()

    def set_stream(self, cuda_stream: int):
        self._ck(lib().psi_b200_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_option(self, name: str, value: int):
        self._ck(lib().psi_b200_set_option(self._h, name.encode(), value))

    def sync(self):
        self._ck(lib().psi_b200_sync(self._h))

    def set_graph(self, g: Graph, ids="internal"):
        node_id = g.internal_id if ids == "internal" else g.coord_id
        self._keep = (g.seq_start, g.seq, g.row_ptr, g.col, node_id)
        self._ck(lib().psi_b200_set_graph(self._h, g.n_nodes, _ptr(g.seq_start), _ptr(g.seq), _ptr(g.row_ptr),
                                          _ptr(g.col), _ptr(node_id)))

    def set_paths(self, p: PathSet):
        self._ck(lib().psi_b200_set_paths(self._h, p.n_paths, _ptr(p.path_ptr), _ptr(p.nodes), _ptr(p.head_off),
                                          _ptr(p.tail_trim)))

    def find_loci(self, step=1) -> int:
        n = C.c_uint64()
        self._ck(lib().psi_b200_find_loci(self._h, step, C.byref(n)))
        return n.value

    def get_loci(self):
        n = C.c_uint64()
        self._ck(lib().psi_b200_get_loci(self._h, None, None, 0, C.byref(n)))
        node = np.zeros(n.value, np.uint32)
        off = np.zeros(n.value, np.uint32)
        if n.value:
            self._ck(lib().psi_b200_get_loci(self._h, _ptr(node), _ptr(off), n.value, C.byref(n)))
        return node, off

    def set_loci(self, node, off):
        node = np.ascontiguousarray(node, np.uint32)
        off = np.ascontiguousarray(off, np.uint32)
        self._ck(lib().psi_b200_set_loci(self._h, len(node), _ptr(node), _ptr(off)))

    def submit_chunk(self, read_ptr, bases, first_read_id=0, distance=0):
        read_ptr = np.ascontiguousarray(read_ptr, np.uint64)
        bases = np.ascontiguousarray(bases, np.uint8)
        self._chunk_keep = (read_ptr, bases)
        self._ck(lib().psi_b200_submit_chunk(self._h, len(read_ptr) - 1, _ptr(read_ptr), _ptr(bases), first_read_id,
                                             distance))

    def submit_chunk_ptr(self, n_reads, read_ptr_addr, bases_addr, first_read_id=0, distance=0):
        """Host pointers given as integers (e.g. pinned torch tensors)."""
        self._ck(lib().psi_b200_submit_chunk(self._h, n_reads, C.c_void_p(read_ptr_addr), C.c_void_p(bases_addr),
                                             first_read_id, distance))

    def submit_chunk_device(self, n_reads, d_read_ptr, d_bases, n_bases, first_read_id=0, distance=0):
        self._ck(lib().psi_b200_submit_chunk_device(self._h, n_reads, C.c_void_p(d_read_ptr), C.c_void_p(d_bases),
                                                    n_bases, first_read_id, distance))

    def submit_chunk_packed(self, p: Packed, distance=0, with_read_ptr=None):
        self._chunk_keep = p
        v = p.struct(with_read_ptr)
        self._ck(lib().psi_b200_submit_chunk_packed(self._h, C.byref(v), distance, 0))

    def submit_chunk_packed_raw(self, n_reads, n_bases, read_len, words_addr, first_read_id=0, distance=0, on_device=False,
                                read_ptr_addr=None, exc_addr=None, n_exc=0):
        """Pointers given as integers (pinned torch tensors / device tensors)."""
        v = PackedChunk()
        v.n_reads, v.first_read_id, v.n_bases, v.read_len = n_reads, first_read_id, n_bases, read_len
        v.read_ptr, v.words, v.exc, v.n_exc = read_ptr_addr, words_addr, exc_addr, n_exc
        self._ck(lib().psi_b200_submit_chunk_packed(self._h, C.byref(v), distance, int(on_device)))

    def seeds_all_async(self, flags=ALL):
        self._ck(lib().psi_b200_seeds_all_async(self._h, flags))

    def wait(self) -> int:
        n = C.c_uint64()
        self._ck(lib().psi_b200_wait(self._h, C.byref(n)))
        return n.value

    def dense_counts(self):
        ns, ne = C.c_uint64(), C.c_uint64()
        self._ck(lib().psi_b200_dense_counts(self._h, C.byref(ns), C.byref(ne)))
        return ns.value, ne.value

    def dense_off_bytes(self) -> int:
        b = C.c_uint()
        self._ck(lib().psi_b200_dense_layout(self._h, C.byref(b)))
        return b.value

    def fetch_dense(self):
        """After seeds_all(flags | DENSE): (dense (n_seeds, 2) u32 {node id, node offset | off-path << 31}, extra (n_extra, 4) u32).
        The library delivers two planes (ids u32, offsets u16 or u32); they are widened into pairs here."""
        ns, ne = self.dense_counts()
        ob = self.dense_off_bytes()
        raw = np.zeros(ns * (4 + ob), np.uint8)
        extra = np.zeros((ne, 4), np.uint32)
        self._ck(lib().psi_b200_fetch_dense(self._h, _ptr(raw), ns, _ptr(extra), ne, C.byref(C.c_uint64()), C.byref(C.c_uint64())))
        return dense_planes(raw, ns, ob), extra

    def dense5_layout(self):
        """(off_bits, available) of PSI_B200_DENSE5 for the current index."""
        ob, av = C.c_uint(), C.c_int()
        self._ck(lib().psi_b200_dense5_layout(self._h, C.byref(ob), C.byref(av)))
        return ob.value, bool(av.value)

    def fetch_dense5(self):
        """After seeds_all(flags | DENSE5): the 5-byte planes decoded into the pairs fetch_dense returns
        (dense (n_seeds, 2) u32 {node id, node offset | off-path << 31}, extra (n_extra, 4) u32)."""
        ns, ne = self.dense_counts()
        raw = np.zeros(ns * 5, np.uint8)
        extra = np.zeros((ne, 4), np.uint32)
        self._ck(lib().psi_b200_fetch_dense(self._h, _ptr(raw), ns, _ptr(extra), ne, C.byref(C.c_uint64()), C.byref(C.c_uint64())))
        return dense5_planes(raw, ns, self.dense5_layout()[0]), extra

    def fetch_dense_async(self, dense_addr: int, cap_seeds: int, extra_addr: int, cap_extra: int):
        self._ck(lib().psi_b200_fetch_dense_async(self._h, C.c_void_p(dense_addr), cap_seeds, C.c_void_p(extra_addr), cap_extra))

    def build_mem_index(self, p: PathSet):
        self._ck(lib().psi_b200_build_mem_index(self._h, p.n_paths, _ptr(p.path_ptr), _ptr(p.nodes), _ptr(p.head_off), _ptr(p.tail_trim)))

    def find_mems(self, max_mem=0) -> np.ndarray:
        """MEM mode over the submitted chunk: (n, 6) u64 {node_id, node_off, read_id, read_off, match_len, gocc}."""
        n = C.c_uint64()
        self._ck(lib().psi_b200_find_mems(self._h, max_mem, C.byref(n)))
        out = np.zeros((n.value, 6), np.uint64)
        if n.value:
            self._ck(lib().psi_b200_fetch_mems(self._h, _ptr(out), n.value, C.byref(n)))
        return out

    def create_distance_index(self, dmin: int, dmax: int):
        self._ck(lib().psi_b200_create_distance_index(self._h, dmin, dmax))

    def verify_distance(self, pairs) -> np.ndarray:
        """pairs: (n, 4) u32 {rank of v, offset, rank of u, offset}; returns n booleans."""
        q = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 4)
        out = np.zeros(len(q), np.uint8)
        self._ck(lib().psi_b200_verify_distance(self._h, len(q), _ptr(q), _ptr(out), 0))
        return out.astype(bool)

    def verify_distance_device(self, n: int, d_pairs: int, d_out: int):
        self._ck(lib().psi_b200_verify_distance(self._h, n, C.c_void_p(d_pairs), C.c_void_p(d_out), 1))

    def seeds_all(self, flags=ALL) -> int:
        n = C.c_uint64()
        self._ck(lib().psi_b200_seeds_all(self._h, flags, C.byref(n)))
        return n.value

    def fetch(self, n=None) -> np.ndarray:
        """(n, 4) u64 records {node_id, node_offset, read_id, read_offset} (reference src/psikt.cpp:172-181)."""
        cnt = C.c_uint64()
        self._ck(lib().psi_b200_fetch(self._h, None, 0, C.byref(cnt)))
        n = cnt.value if n is None else min(n, cnt.value)
        out = np.zeros((n, 4), np.uint64)
        if n:
            self._ck(lib().psi_b200_fetch(self._h, _ptr(out), n, C.byref(cnt)))
        return out

    def fetch32(self, n=None) -> np.ndarray:
        """(n, 4) u32 records, same fields, after seeds_all(flags | COMPACT)."""
        cnt = C.c_uint64()
        self._ck(lib().psi_b200_fetch32(self._h, None, 0, C.byref(cnt)))
        n = cnt.value if n is None else min(n, cnt.value)
        out = np.zeros((n, 4), np.uint32)
        if n:
            self._ck(lib().psi_b200_fetch32(self._h, _ptr(out), n, C.byref(cnt)))
        return out

    def fetch32_into(self, addr: int, cap: int) -> int:
        cnt = C.c_uint64()
        self._ck(lib().psi_b200_fetch32(self._h, C.c_void_p(addr), cap, C.byref(cnt)))
        return cnt.value

    def fetch_kinds(self) -> np.ndarray:
        """Per record: 1 = on an indexed path, 2 = off-path (same order as fetch())."""
        cnt = C.c_uint64()
        self._ck(lib().psi_b200_fetch_kinds(self._h, None, 0, C.byref(cnt)))
        out = np.zeros(cnt.value, np.uint8)
        if cnt.value:
            self._ck(lib().psi_b200_fetch_kinds(self._h, _ptr(out), cnt.value, C.byref(cnt)))
        return out

    def fetch_into(self, addr: int, cap: int) -> int:
        cnt = C.c_uint64()
        self._ck(lib().psi_b200_fetch(self._h, C.c_void_p(addr), cap, C.byref(cnt)))
        return cnt.value

    def fetch_device(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(lib().psi_b200_fetch_device(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def counters(self) -> dict:
        c = Counters()
        self._ck(lib().psi_b200_counters(self._h, C.byref(c)))
        return c.as_dict()

    def reset_counters(self):
        self._ck(lib().psi_b200_reset_counters(self._h))


def dense_planes(raw: np.ndarray, n_seeds: int, off_bytes: int) -> np.ndarray:
    """ids[n_seeds] u32 + offs[n_seeds] u16/u32 (top bit = off-path) -> (n_seeds, 2) u32 {node id, offset | off-path << 31}."""
    raw = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
    ids = raw[: 4 * n_seeds].view(np.uint32)
    if off_bytes == 2:
        o = raw[4 * n_seeds: 6 * n_seeds].view(np.uint16).astype(np.uint32)
        offs = (o & 0x7FFF) | ((o >> 15) << 31)
    else:
        offs = raw[4 * n_seeds: 8 * n_seeds].view(np.uint32)
    return np.column_stack([ids, offs]).astype(np.uint32)


def dense5_planes(raw: np.ndarray, n_seeds: int, off_bits: int) -> np.ndarray:
    """PSI_B200_DENSE5 planes (u32 low words, then one byte per seed) -> (n_seeds, 2) u32 {node id, offset | off-path << 31},
    NIL32 / 0 for a seed without a hit: the representation dense_planes gives for PSI_B200_DENSE."""
    lo = raw[:4 * n_seeds].view(np.uint32).astype(np.uint64)
    hi = raw[4 * n_seeds:5 * n_seeds].astype(np.uint64)
    nil = (lo == NIL32) & (hi == 0xFF)
    e = lo | ((hi & np.uint64(0x7F)) << np.uint64(32))
    out = np.zeros((n_seeds, 2), np.uint32)
    out[:, 0] = np.where(nil, NIL32, e >> np.uint64(off_bits)).astype(np.uint32)
    out[:, 1] = np.where(nil, 0, (e & np.uint64((1 << off_bits) - 1)) | ((hi >> np.uint64(7)) << np.uint64(31))).astype(np.uint32)
    return out


def seed_layout(read_ptr, k: int, d: int):
    """(read index, read offset) of every seed of a chunk in seed order (sequence.hpp:1712): what the position in
    the dense result array stands for."""
    lens = np.diff(np.asarray(read_ptr, np.int64))
    d = d or k
    cnt = np.where(lens >= k, (lens - k) // d + 1, 0)
    read = np.repeat(np.arange(len(lens), dtype=np.int64), cnt)
    first = np.concatenate([[0], np.cumsum(cnt)])[:-1]
    off = (np.arange(int(cnt.sum()), dtype=np.int64) - np.repeat(first, cnt)) * d
    return read, off


def dense_to_records(dense: np.ndarray, extra: np.ndarray, read_ptr, k: int, d: int, first_read_id=0):
    """Dense per-seed results + extra list -> ((n, 4) u64 records {node_id, node_off, read_id, read_off}, kinds)."""
    read, off = seed_layout(read_ptr, k, d)
    assert len(read) == len(dense), (len(read), len(dense))
    hit = dense[:, 0] != NIL32
    rec = np.column_stack([dense[hit, 0].astype(np.uint64), (dense[hit, 1] & 0x7FFFFFFF).astype(np.uint64),
                           (read[hit] + first_read_id).astype(np.uint64), off[hit].astype(np.uint64)])
    kinds = np.where(dense[hit, 1] >> 31, 2, 1).astype(np.uint8)
    if len(extra):
        ex = np.column_stack([extra[:, 0].astype(np.uint64), extra[:, 1].astype(np.uint64), extra[:, 2].astype(np.uint64),
                              (extra[:, 3] & 0x7FFFFFFF).astype(np.uint64)])
        rec = np.concatenate([rec, ex])
        kinds = np.concatenate([kinds, np.where(extra[:, 3] >> 31, 2, 1).astype(np.uint8)])
    return rec.reshape(-1, 4), kinds


def canonical(records: np.ndarray) -> np.ndarray:
    """Reference-CLI records {node, node_off, read, read_off} -> the canonical seed
    set: unique rows (read_id, read_offset, node_id, node_offset), sorted."""
    if records.size == 0:
        return np.zeros((0, 4), np.uint64)
    t = np.ascontiguousarray(records.reshape(-1, 4)[:, [2, 3, 0, 1]])
    return np.unique(t, axis=0)
