"""N > 1 host logic on CPU: read sharding, disjoint global read ids, and the counts/histogram all-reduce
over a world_size-2 gloo group.  The per-shard seed sets come from the oracle here (no GPU); the GPU
path uses exactly the same sharding in bench.py."""
import os
import socket

import numpy as np
import pytest

import util
from psi_b200 import shard


def test_shard_bounds_partition_and_balance():
    rng = np.random.default_rng(1)
    lens = rng.integers(50, 250, size=10007)
    read_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    for world in (1, 2, 3, 4, 8):
        b = shard.shard_bounds(read_ptr, world)
        assert b[0] == 0 and b[-1] == len(lens) and np.all(np.diff(b) >= 0) and len(b) == world + 1
        per = [int(read_ptr[b[i + 1]] - read_ptr[b[i]]) for i in range(world)]
        assert max(per) - min(per) <= 2 * 250
    # more ranks than reads: empty shards, still a partition
    b = shard.shard_bounds(np.array([0, 10, 20], np.uint64), 8)
    assert b[0] == 0 and b[-1] == 2 and np.all(np.diff(b) >= 0)
    # shard_of keeps global read ids
    bases = np.zeros(int(read_ptr[-1]), np.uint8)
    got = [shard.shard_of(read_ptr, bases, r, 4, first_read_id=100) for r in range(4)]
    assert got[0][2] == 100 and sum(len(g[0]) - 1 for g in got) == len(lens)
    assert sum(len(g[1]) for g in got) == len(bases)
    for g in got:
        assert g[0][0] == 0 and g[0][-1] == len(g[1])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from oracle import oracle_py as orc
    from psi_b200 import capi
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = {x["name"]: x for x in util.golden_index()["cases"]}["x_k12"]
    g = capi.Graph.load_gfa(util.GOLDEN / c["gfa"])
    rp, bases = util.read_fasta(util.GOLDEN / c["reads"])
    sp, sb, first = shard.shard_of(rp, bases, rank, world)
    tuples, _ = orc.seeds_closed_form(orc.OGraph.of(g), orc.OReads(sp, sb, first), c["k"], c["d"])
    hist = shard.hits_per_read_histogram(tuples[:, 0], len(sp) - 1, first)
    counts, total_hist = shard.all_reduce_counts({"reads": len(sp) - 1, "hits": len(tuples)}, hist)
    np.save(os.path.join(out_dir, f"set_{rank}.npy"), tuples)
    if rank == 0:
        np.save(os.path.join(out_dir, "hist.npy"), total_hist)
        np.save(os.path.join(out_dir, "counts.npy"), np.array([counts["reads"], counts["hits"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_run_equals_the_whole(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), os.fspath(tmp_path)), nprocs=world, join=True)
    c = {x["name"]: x for x in util.golden_index()["cases"]}["x_k12"]
    parts = [np.load(tmp_path / f"set_{r}.npy") for r in range(world)]
    # read ids of the shards are disjoint and ascending by rank: concatenation is already canonical
    assert parts[0][:, 0].max() < parts[1][:, 0].min()
    whole = np.concatenate(parts)
    assert len(whole) == c["count"] and util.md5_tuples(whole) == c["md5"]
    reads, hits = np.load(tmp_path / "counts.npy")
    assert reads == 10000 and hits == c["count"]
    hist = np.load(tmp_path / "hist.npy")
    assert hist.sum() == 10000 and (hist * np.arange(len(hist))).sum() == c["count"]
