#!/usr/bin/env python
"""bench.py -- reads/s and seeds/s of fully-sensitive seed finding on the
synthetic chr22-shape graph (BASELINE.json configs[1]), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]    # the reference's own CPU path (oracle/_ref)

A step = one pass of the hot path (seeding -> seeds_on_paths / seeds_off_paths ->
seed records; one fused kernel in index mode) over one batch of 1 M synthetic 100 bp reads.
  value : whole-job reads/s with the batch already resident in HBM
  e2e   : same through the C-ABI with pinned HOST buffers, H2D of the reads and D2H of
          the seed records inside the timed region
  roofline : the dominant kernel (fused one-pass kernel, else the seeds_on_paths probe), algorithmic bytes /
             CUDA-event time vs measured HBM peak; the probe kernel alone is reported beside it
  cpu_baseline : the unmodified reference (oracle/_ref/psi_ref_driver) on the host cores, bounded sample
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, os.fspath(ROOT))

# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of one launch of the dominant kernel on this workload from
# the last `ncu --set full` capture (see profiles/); None until a capture of the current kernel exists.
TRAFFIC = {
    "seeds_on_paths_kernel": (662.9e6, "profiles/r01i_kernels_ncu_raw.csv: seeds_on_paths_kernel<8>, dram__bytes_read.sum 635.3 MB + "
                                       "dram__bytes_write.sum 27.6 MB (mean of 2 launches)"),
    "seeds_fused_kernel_dense_packed": (645.3e6, "profiles/r02c_fused_ncu_raw.csv: seeds_fused_kernel<8, 1, 4, dense, packed>, "
                                                 "dram__bytes_read.sum 614.0 MB + dram__bytes_write.sum 31.3 MB (mean of 3 launches; "
                                                 "below the line accounting because 40 % of the sectors hit L2)"),
}

K = 20
READ_LEN = 100
READS_PER_BATCH = 1_000_000
N_PATHS = 16
N_BATCHES = 3


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------- clocks --

class ClockSampler:
    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.max_mhz = None
        self.ready = threading.Event()     # nvidia-smi has attached to the driver and delivered its first sample
        self.t_from = 0.0

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))
            self.ready.set()

    def mark(self):
        """Samples from here on count: nvidia-smi is started early (attaching to the driver takes it a second or two on an
        8-GPU box, and doing that beside a 3 ms timed loop perturbs the loop), its idle-time samples are dropped."""
        self.ready.wait(10.0)
        self.t_from = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if t < self.t_from:
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------- workload --

def build_graph(shape: str):
    from bench_support import synth
    from psi_b200 import capi
    spec = synth.SHAPES[shape]
    a = synth.graph_arrays_multi(spec["components"]) if "components" in spec else synth.graph_arrays(**spec)
    return capi.Graph.from_arrays(a["ids"], a["seq_start"], a["seq"], a["row_ptr"], a["col"], a["path_ptr"],
                                  a["path_nodes"], sort=True)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------ reference (CPU) --

def reference_run(n_reads_per_core=None, cores=None, shape="chr22_1_51", budget_s=8.0):
    """The unmodified reference on the host cores: one psi_ref_driver process per core on
    disjoint read ranges of one FASTA (the reference seed path is single-threaded).
    Returns dict(value reads/s, cores, sample, seeds_per_s, seconds)."""
    from bench_support import synth
    from oracle import oracle_py as orc
    if not orc.have_reference():
        raise RuntimeError("oracle/_ref/psi_ref_driver missing")
    cores = cores or max(1, (os.cpu_count() or 1))
    cores = min(cores, 32)
    g = build_graph(shape)
    # calibrated so that one process takes roughly budget_s: ~6.7e3 reads/s/core measured in the survey on this shape
    n_per = n_reads_per_core or max(2000, int(2500 * budget_s))
    n_reads = n_per * cores
    rp, bases = synth.reads(g, n_reads, READ_LEN, 1002)
    td = tempfile.mkdtemp(prefix="psi_ref_")
    gfa = os.path.join(td, "g.gfa")
    fa = os.path.join(td, "r.fa")
    g.write_gfa(gfa)
    with open(fa, "wb") as f:
        b = bases.reshape(n_reads, READ_LEN)
        for i in range(n_reads):
            f.write(b">r%d\n" % i)
            f.write(b[i].tobytes())
            f.write(b"\n")
    env = dict(os.environ, TMPDIR=td, OMP_PROC_BIND="false", OMP_NUM_THREADS="1")
    stats, used, last_err, tries = [], cores, "", 0
    while used >= 1:
        # one single-threaded reference process per core on disjoint read ranges; halve the count if a process dies
        # (e.g. out of memory on a box with many cores)
        procs = []
        for c in range(used):
            cmd = [os.fspath(orc.REF_DRIVER), "--gfa", gfa, "--fastq", fa, "-k", str(K), "-d", str(K), "-n", str(N_PATHS),
                   "--first-read", str(c * n_per), "--max-reads", str(n_per)]
            procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
        stats = []
        for p in procs:
            out, err = p.communicate()
            lines = out.strip().splitlines()
            if p.returncode == 0 and lines:
                stats.append(json.loads(lines[-1]))
            else:
                last_err = (err or "").strip().splitlines()[-1:] or [f"exit code {p.returncode}"]
        if len(stats) == used:
            break
        # The reference's TraverserBFS::advance pushes a copy of a state onto the vector the state lives in and keeps using
        # the reference afterwards (traverser_bfs.hpp:147-158): a heap use-after-free whenever the vector grows at a node
        # with three or more successors (AddressSanitizer report: profiles/r02_reference_asan_use_after_free.txt).  Whether
        # it crashes depends on the heap and on the randomly picked paths, so one more try at the same process count,
        # then halve (a process may also have run out of memory on a box with many cores)
        tries += 1
        nxt = used if tries % 2 == 1 else used // 2
        log(f"[bench] reference: {used - len(stats)} of {used} processes failed ({last_err}); retrying with {nxt}")
        used = nxt
    if not stats or len(stats) != used:
        raise RuntimeError(f"reference driver failed: {last_err}")
    cores = used
    import shutil
    shutil.rmtree(td, ignore_errors=True)
    # timed region = seeding + seeds_on_paths + seeds_off_paths (index build excluded, as for the GPU arm)
    t = max(s["t_seeding"] + s["t_on"] + s["t_off"] for s in stats)
    reads = sum(s["reads"] for s in stats)
    hits = sum(s["raw_on"] + s["raw_off"] for s in stats)
    return {"value": reads / t, "unit": "reads/s", "cores": cores, "kind": "reference",
            "sample": f"synthetic chr22-shape graph at 1/51 scale (1 Mbp backbone, 19 608 sites, {N_PATHS} paths), "
                      f"{reads} x {READ_LEN} bp reads, k={K}: {cores} single-threaded reference processes on disjoint "
                      f"read ranges; timed = seeding + seeds_on_paths + seeds_off_paths, max over processes",
            "seconds": t, "raw_hits_per_s": hits / t, "t_index_s": max(s["t_index"] for s in stats)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the inputs (synthetic graph -> GFA, reads -> FASTA) are built with the host half of the library only: this process
    # never maps the CUDA library
    from psi_b200 import capi
    capi.use_host_library()
    try:
        vals = []
        for i in range(args.warmup + args.steps):
            r = reference_run(budget_s=4.0)
            if i >= args.warmup:
                vals.append(r)
        v = float(np.mean([x["value"] for x in vals]))
        ms = float(np.mean([x["seconds"] for x in vals])) * 1e3
        last = vals[-1]
        line = {"impl": "reference", "metric": "reads/s (fully-sensitive seed finding, chr22-shape graph, k=20)",
                "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64",
                "data": "synthetic", "config": {"workload": "chr22-shape graph at 1/51 scale, bounded sample per step"},
                "cpu_baseline": {"value": v, "unit": "reads/s", "cores": last["cores"], "kind": "reference",
                                 "sample": last["sample"]},
                "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    except Exception as e:  # the reference binary did not travel / cannot run
        line = {"impl": "reference", "unavailable": f"{type(e).__name__}: {e}"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------ GPU arm --

class Workload:
    """One graph + index + read batches on this rank's GPU: the pipelines (a context and its forks sharing the resident
    index), the batches as ASCII and as 2-bit words, resident in HBM and in pinned host memory."""

    def __init__(self, torch, dev, rank, shape, k, read_len, n_reads, n_batches, n_paths, n_pipes, offpath_mode=0, opts=()):
        from bench_support import synth
        from psi_b200 import capi
        self.torch, self.dev, self.rank, self.k, self.read_len, self.n_reads = torch, dev, rank, k, read_len, n_reads
        t0 = time.time()
        self.g = g = build_graph(shape)
        self.ps = ps = g.pick_paths(n_paths, seed=1)
        self.ctx = ctx = capi.Context(k, dev.index)
        ctx.set_option("offpath_mode", offpath_mode)
        for kv in opts:
            name, val = kv.split("=")
            ctx.set_option(name, int(val))
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx.set_graph(g, ids="internal")
        ctx.set_paths(ps)
        self.n_loci = ctx.find_loci()
        self.c0 = c0 = ctx.counters()
        if rank == 0:
            log(f"[bench] {shape}: graph {g.n_nodes} nodes / {g.n_bases} bp; index {c0['n_index_kmers']} k-mers, "
                f"{c0['index_bytes'] / 1e6:.0f} MB, slot {c0['index_slot_bytes']} B, locus codes by "
                f"{'rank' if c0['code_by_rank'] else 'id'} ({c0['code_off_bits']} offset bits), build {c0['ms_index_build']:.0f} ms; "
                f"{self.n_loci} starting loci, {c0['n_offpath_walks']} uncovered walks -> {c0['n_offpath_entries']} off-path entries "
                f"(mode {c0['offpath_mode']}) in {c0['ms_find_loci']:.0f} ms; setup {time.time() - t0:.1f} s")
        self.per_read = (read_len - k) // k + 1
        self.n_seeds = n_reads * self.per_read
        # distinct read batches per rank (weak scaling: every GPU processes its own shard of the read set)
        self.ascii_h, self.ascii_d, self.words_h, self.words_d = [], [], [], []
        t_pack = 0.0
        for b in range(n_batches):
            rp, bases = synth.reads(g, n_reads, read_len, 1002 + 1000 * rank + b)
            hp = torch.from_numpy(rp.view(np.int64)).pin_memory()
            hb = torch.from_numpy(bases).pin_memory()
            t1 = time.perf_counter()
            pk = capi.Packed.pack(rp, bases, rank * n_reads)       # what psi_b200_reader_next_packed does per chunk
            t_pack += time.perf_counter() - t1
            assert pk.read_len == read_len and len(pk.exc) == 0
            hw = torch.from_numpy(pk.words.view(np.int64)).pin_memory()
            self.ascii_h.append((hp, hb))
            self.ascii_d.append((hp.to(dev), hb.to(dev)))
            self.words_h.append(hw)
            self.words_d.append(hw.to(dev))
        self.pack_gbs = n_batches * n_reads * read_len / t_pack / 1e9 / 2   # pack_bases is run twice by Packed.pack (count, then fill)
        self.n_batches = n_batches
        self.pipes = [ctx] + [ctx.fork() for _ in range(max(n_pipes, 2) - 1)]     # the ASCII comparison arm drives two
        n_extra_cap = self.n_seeds // 8 + 4096
        self.dense_h = [torch.empty((self.n_seeds, 2), dtype=torch.int32).pin_memory() for _ in self.pipes]
        self.extra_h = [torch.empty((n_extra_cap, 4), dtype=torch.int32).pin_memory() for _ in self.pipes]
        self.rec_h = [None] * len(self.pipes)
        torch.cuda.synchronize()

    # ---- one step on pipeline p: queue the chunk, the kernels and (e2e) the copy of the results ----
    def submit(self, p, i, fmt, where):
        cx, b, first = self.pipes[p], i % self.n_batches, self.rank * self.n_reads
        if fmt == "packed":
            w = (self.words_d if where == "device" else self.words_h)[b]
            cx.submit_chunk_packed_raw(self.n_reads, self.n_reads * self.read_len, self.read_len, w.data_ptr(), first, self.k,
                                       on_device=(where == "device"))
        elif where == "device":
            dp, db = self.ascii_d[b]
            cx.submit_chunk_device(self.n_reads, dp.data_ptr(), db.data_ptr(), db.numel(), first, self.k)
        else:
            hp, hb = self.ascii_h[b]
            cx.submit_chunk_ptr(self.n_reads, hp.data_ptr(), hb.data_ptr(), first, self.k)

    def close(self):
        for cx in self.pipes[1:]:
            cx.close()
        self.ctx.close()


def main_gpu(args):
    import torch
    import torch.distributed as dist
    from psi_b200 import capi, shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def run_async(W, steps, warmup, n_pipes, fmt, where, flags, fetch):
        """K steps through the C-ABI issued by ONE host thread round-robin over n_pipes contexts: submit_chunk* ->
        seeds_all_async (-> fetch_dense_async); a context's previous step is completed (psi_b200_wait) only when the
        context comes round again, so n_pipes chunks are in flight.  Timed with CUDA events after a barrier +
        synchronize; every step has been waited for before the closing event.  Returns (ms, hits, launches, fused kernel
        ms per step from the contexts' CUDA events)."""
        pipes = W.pipes[:n_pipes]
        busy = [False] * n_pipes

        def issue(i):
            p = i % n_pipes
            n = 0
            if busy[p]:
                n = pipes[p].wait()
            W.submit(p, i, fmt, where)
            pipes[p].seeds_all_async(flags)
            if fetch:
                pipes[p].fetch_dense_async(W.dense_h[p].data_ptr(), W.n_seeds, W.extra_h[p].data_ptr(), W.extra_h[p].shape[0])
            busy[p] = True
            return n

        def drain():
            n = 0
            for p in range(n_pipes):
                if busy[p]:
                    n += pipes[p].wait()
                    busy[p] = False
            return n

        for i in range(max(warmup, n_pipes)):
            issue(i)
        drain()
        barrier()
        for cx in pipes:
            cx.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hits = 0
        e0.record()
        for i in range(steps):
            hits += issue(warmup + i)
        hits += drain()
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        cs = [cx.counters() for cx in pipes]
        launches = sum(c["launches"] for c in cs)
        timed = sum(c["timed_steps"] for c in cs)
        k_ms = sum(c["ms_probe_sum"] for c in cs) / timed if timed else 0.0
        return ms, hits, launches, k_ms

    def run_sync(W, steps, warmup, fmt, flags):
        """The same K steps on ONE context, each completed before the next is issued."""
        cx = W.pipes[0]
        for i in range(warmup):
            W.submit(0, i, fmt, "device")
            cx.seeds_all(flags)
        barrier()
        cx.reset_counters()
        acc = {"ms_on": 0.0, "ms_probe": 0.0, "ms_pack": 0.0, "ms_resolve": 0.0, "n_hits": 0, "n_hits_on": 0, "n_seeds": 0,
               "n_on_probe_sectors": 0, "n_walks": 0, "ms_off": 0.0, "ms_read_index": 0.0}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hits = 0
        e0.record()
        for i in range(steps):
            W.submit(0, warmup + i, fmt, "device")
            hits += cx.seeds_all(flags)
            c = cx.counters()
            for k_ in acc:
                acc[k_] += c[k_]
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), hits, acc, cx.counters()["launches"]

    def run_threads_ascii(W, steps, warmup, n_pipes, compact=True):
        """Round 1's e2e arm, kept for comparison: ASCII chunk up, 16-byte records per hit down, one host thread per
        pipeline, every call synchronous."""
        def step(i, p):
            cx = W.pipes[p]
            if W.rec_h[p] is None:
                W.rec_h[p] = torch.empty((2 * W.n_seeds, 4), dtype=torch.int32 if compact else torch.int64).pin_memory()
            W.submit(p, i, "ascii", "host")
            if compact:
                cx.seeds_all(capi.ALL | capi.COMPACT)
                return cx.fetch32_into(W.rec_h[p].data_ptr(), W.rec_h[p].shape[0])
            cx.seeds_all(capi.ALL)
            return cx.fetch_into(W.rec_h[p].data_ptr(), W.rec_h[p].shape[0])
        for i in range(max(warmup, n_pipes)):
            step(i, i % n_pipes)
        barrier()
        hits, errors = [0] * n_pipes, []

        def work(p):
            try:
                for i in range(p, steps, n_pipes):
                    hits[p] += step(warmup + i, p)
            except Exception as e:   # surfaced after join
                errors.append(e)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        threads = [threading.Thread(target=work, args=(p,)) for p in range(n_pipes)]
        e0.record()
        for t in threads:
            t.start()
        for t in threads:
            t.join()          # every step ended with a synchronous fetch: all device work is done
        e1.record()
        barrier()
        if errors:
            raise errors[0]
        return max_over_ranks(e0.elapsed_time(e1)), sum(hits)

    n_reads, steps, warmup = args.reads, args.steps, args.warmup
    n_pipes = max(1, args.pipelines)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()        # polls every 100 ms from now on; only the samples taken during the timed loops are kept
    W = Workload(torch, dev, rank, args.shape, K, READ_LEN, n_reads, N_BATCHES, N_PATHS, n_pipes, args.offpath_mode, args.opt)
    ctx, c0 = W.ctx, W.c0
    DENSE = capi.ALL | capi.DENSE
    # the results' wire format: 5 bytes per seed (PSI_B200_DENSE5) when the graph's loci fit 39 bits, else 6 / 8
    d5_off_bits, d5 = ctx.dense5_layout()
    FAST = capi.ALL | (capi.DENSE5 if d5 else capi.DENSE)
    res_bytes = 5 if d5 else 4 + ctx.dense_off_bytes()
    if rank == 0:
        sampler.mark()
    barrier()
    # value: 2-bit chunks resident in HBM -> dense results resident in HBM, n_pipes chunks in flight from one host thread
    ms_val, hits_val, launches_val, kms_val = run_async(W, steps, warmup, n_pipes, "packed", "device", FAST, False)
    # e2e: the same call sequence with pinned HOST buffers: 2-bit chunk up, dense results down, inside the timed region
    ms_e2e, hits_e2e, launches_e2e, kms_e2e = run_async(W, steps, warmup, n_pipes, "packed", "host", FAST, True)
    # the same end-to-end loop with round 2's first result format (6 bytes per seed: u32 node id + u16 offset)
    ms_e2e6, hits_e2e6, _, _ = run_async(W, steps, warmup, n_pipes, "packed", "host", DENSE, True) if d5 else (ms_e2e, hits_e2e, 0, 0)
    assert hits_e2e6 == hits_e2e
    # beside them: ASCII chunks (the packing then runs inside the fused kernel) and round 1's record formats
    ms_val_ascii, hits_a, _, kms_ascii = run_async(W, steps, warmup, n_pipes, "ascii", "device", DENSE, False)
    ms_val_rec, hits_r, _, kms_rec = run_async(W, steps, warmup, n_pipes, "ascii", "device", capi.ALL, False)
    ms_e2e_ascii, hits_ea = run_threads_ascii(W, steps, warmup, 2, compact=True)
    ms_one, hits_one, acc_one, launches_one = run_sync(W, steps, warmup, "packed", DENSE)
    assert hits_val == hits_e2e == hits_a == hits_r == hits_ea == hits_one, (hits_val, hits_e2e, hits_a, hits_r, hits_ea, hits_one)
    # the same resident step through the separate seeding / probe / resolve kernels: the seeds_on_paths probe alone
    ctx.set_option("fused", 0)
    ms_sep, hits_sep, acc_sep, _ = run_sync(W, steps, warmup, "ascii", capi.ALL)
    ctx.set_option("fused", 1)
    assert hits_sep == hits_val, (hits_sep, hits_val)
    clocks = sampler.stop() if rank == 0 else None      # sampled over all the timed loops above (value, e2e and the comparison arms)

    # per-shard counts and a hits-per-read histogram, reduced with NCCL (the only collective on this path)
    W.submit(0, 0, "packed", "device")
    ctx.seeds_all(DENSE)
    dense, extra = ctx.fetch_dense()
    per_read = (dense[:, 0] != capi.NIL32).reshape(n_reads, W.per_read).sum(axis=1)
    if len(extra):
        per_read += np.bincount(extra[:, 2].astype(np.int64) - rank * n_reads, minlength=n_reads)[:n_reads]
    hist = np.bincount(np.minimum(per_read, shard.HIST_BINS - 1), minlength=shard.HIST_BINS)
    counts, hist = shard.all_reduce_counts({"reads": n_reads * steps, "seeds": acc_one["n_seeds"], "hits": hits_val,
                                            "hits_on": acc_one["n_hits_on"]}, hist, device=dev)

    barrier()      # every rank copies at the same moment: the figure is one GPU's share of the box's host bandwidth
    pcie = pcie_ceiling(torch, dev, int(W.words_h[0].numel() * 8), int(res_bytes * W.n_seeds))
    barrier()
    drop_in = psikt_drop_in(W) if (world == 1 and rank == 0 and not args.no_other_configs) else None
    other = {}
    if world == 1 and not args.no_other_configs:
        other = other_configs(torch, dev, run_async, run_sync, capi)

    if rank == 0:
        peak, peak_src = peaks()
        reads_total, seeds_total, hits_total = counts["reads"], counts["seeds"], counts["hits"]
        value = reads_total / (ms_val * 1e-3)
        seeds_step, hits_step = acc_one["n_seeds"] / steps, acc_one["n_hits"] / steps
        words_bytes = int(W.words_h[0].numel() * 8)
        ascii_bytes = int(W.ascii_h[0][1].numel() + W.ascii_h[0][0].numel() * 8)
        off_bytes = res_bytes - 4
        extra_copy = (n_reads * READ_LEN // K + n_reads + 1) // 256 + 256     # speculative share of the extra list copied with every step

        def roof_of(kernel, launch_ms, wall_ms, in_bytes, out_per_seed, traffic_key, what):
            # per query seed 8 B (k-mer) + 32 B (P = 1: one index sector), per hit 8 + 16 B.
            # launch_ms = mean CUDA-event duration of a launch; with several pipelines the launches overlap on the
            # device (concurrency = launch_ms / wall time per launch), so what the GPU achieves is bytes per launch /
            # WALL time per launch = bytes x concurrency / launch_ms.
            alg = seeds_step * 40 + hits_step * 24
            line = in_bytes + seeds_step * 128 + seeds_step * out_per_seed       # what the kernel must move at DRAM-line granularity
            t = min(launch_ms, wall_ms) * 1e-3
            tr = TRAFFIC.get(traffic_key, (None, "no ncu --set full capture of this kernel yet"))
            return {"bound": "hbm", "kernel": kernel, "achieved": alg / t / 1e9 if t > 0 else 0.0, "peak": peak, "unit": "GB/s",
                    "frac": alg / t / 1e9 / peak if t > 0 else 0.0, "traffic": tr[0], "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg, "launch_ms": launch_ms, "wall_ms_per_launch": wall_ms,
                    "concurrency": launch_ms / wall_ms if wall_ms > 0 else 1.0,
                    "accounting": "SURVEY 8(d), P = 1: seeds x (8 B k-mer + 32 B index sector) + hits x (8 B locus + 16 B record); "
                                  "achieved = those bytes / min(launch_ms, wall time per launch)",
                    "line_accounting": {"bytes": what, "bytes_per_launch": line, "achieved": line / t / 1e9 if t > 0 else 0.0,
                                        "frac": line / t / 1e9 / peak if t > 0 else 0.0,
                                        "why": "a random access costs one 128-byte DRAM line whatever part of it is used "
                                               "(profiles/r01c_gather_peak.md): the reachable ceiling is the random-line rate"},
                    "probes_per_s": seeds_step / t if t > 0 else 0.0, "random_line_ceiling_probes_per_s": 4.0e10,
                    "timed_in": "the loop that yields `value` (CUDA events on each pipeline's stream around every launch)",
                    "launch_ms_alone": acc_one["ms_probe"] / steps,
                    "traffic_source": tr[1]}

        roof = roof_of("seeds_fused_kernel<8, 1, 4, dense, packed>", kms_val, ms_val / steps, words_bytes, 4 + off_bytes,
                       "seeds_fused_kernel_dense_packed",
                       f"2-bit chunk in + per seed one 128 B bucket line + {4 + off_bytes} B result out")
        probe_ms = acc_sep["ms_probe"] / steps
        probe_roof = roof_of("seeds_on_paths_kernel<8>", probe_ms, probe_ms, seeds_step * 9, 9, "seeds_on_paths_kernel",
                             "per seed 8 B k-mer + 1 B validity in, one 128 B bucket line, 9 B result out")
        probe_roof["timed_in"] = "the one-context loop of separate_kernels"
        probe_roof.pop("launch_ms_alone")
        line = {
            "metric": f"reads/s (fully-sensitive seed finding, {args.shape}-shape graph, k={K})",
            "value": value, "unit": "reads/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_val / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/u64", "data": "synthetic",
            "config": {"workload": f"synthetic {args.shape}-shape graph ({W.g.n_bases} bp, {W.g.n_nodes} nodes, {N_PATHS} paths, "
                                   f"{W.n_loci} starting loci), {n_reads} x {READ_LEN} bp reads per GPU per step, k={K}, d={K}",
                       "l2": f"{N_BATCHES} distinct read batches cycled and a {c0['index_bytes'] / 1e6:.0f} MB index probed at random: "
                             f"inputs larger than L2",
                       "sharding": "reads sharded by rank, graph + index replicated, NCCL all-reduce of counts only",
                       "formats": "chunk = 2-bit words (psi_b200_packed_chunk, what psi_b200_reader_next_packed hands over; "
                                  f"the reader's packer ran at {W.pack_gbs:.2f} GB/s of characters on one host core here); "
                                  + ("results = 5 bytes per seed: (node id << offset bits | node offset) in 39 bits + the off-path bit "
                                     "(PSI_B200_DENSE5)" if d5 else "results = dense {node id, node offset} per seed (PSI_B200_DENSE)"),
                       "pipelines": f"{n_pipes} contexts per GPU sharing one resident index, all driven by ONE host thread through "
                                    "psi_b200_seeds_all_async / psi_b200_wait"},
            "seeds_per_s": hits_total / (ms_val * 1e-3), "query_seeds_per_s": seeds_total / (ms_val * 1e-3),
            "gpu_launches": int(launches_val), "gpu_launches_e2e": int(launches_e2e),
            "e2e": {"value": reads_total / (ms_e2e * 1e-3), "unit": "reads/s",
                    "h2d_bytes_per_step": words_bytes, "d2h_bytes_per_step": int((4 + off_bytes) * W.n_seeds + 16 * extra_copy + 8 * 12),
                    "ms_per_step": ms_e2e / steps, "pipelines": n_pipes, "host_threads": 1,
                    "fused_kernel_ms": kms_e2e,
                    "pcie": dict(pcie, floor_ms_per_step=max(words_bytes / (pcie["h2d_bidir_gbs"] * 1e6), (4 + off_bytes) * W.n_seeds / (pcie["d2h_bidir_gbs"] * 1e6)),
                                 note="rank 0's link; at N > 1 every rank runs the same copies at the same moment, so this is one GPU's share of the box's host bandwidth"),
                    "formats": f"up: 2-bit words of the chunk (pinned host memory); down: {res_bytes} bytes per seed ("
                               + ("PSI_B200_DENSE5: u32 + u8 planes holding (node id << offset bits | node offset) and the off-path bit"
                                  if d5 else f"u32 node id, u{8 * off_bytes} node offset | off-path bit")
                               + ") + the extra list of multi-locus seeds + the step's counters",
                    "with_6_byte_results": {"value": reads_total / (ms_e2e6 * 1e-3), "ms_per_step": ms_e2e6 / steps,
                                            "d2h_bytes_per_step": int((4 + ctx.dense_off_bytes()) * W.n_seeds + 16 * extra_copy + 8 * 12)}},
            "roofline": roof,
            "value_ascii_chunks": {"value": reads_total / (ms_val_ascii * 1e-3), "unit": "reads/s", "ms_per_step": ms_val_ascii / steps,
                                   "fused_kernel_ms": kms_ascii,
                                   "note": "ASCII chunk resident in HBM (2-bit packing inside the fused kernel), dense results"},
            "value_ascii_chunks_records": {"value": reads_total / (ms_val_rec * 1e-3), "unit": "reads/s", "ms_per_step": ms_val_rec / steps,
                                           "fused_kernel_ms": kms_rec,
                                           "note": "ASCII chunk in, 4 x u64 records per hit out: round 1's `value` configuration"},
            "value_one_pipeline": {"value": reads_total / (ms_one * 1e-3), "unit": "reads/s", "ms_per_step": ms_one / steps,
                                   "fused_kernel_ms": acc_one["ms_probe"] / steps, "step_kernels_ms": acc_one["ms_on"] / steps,
                                   "note": "the same K steps on ONE context, each step waited for before the next is issued"},
            "e2e_ascii_chunks_records": {"value": reads_total / (ms_e2e_ascii * 1e-3), "unit": "reads/s", "ms_per_step": ms_e2e_ascii / steps,
                                         "h2d_bytes_per_step": ascii_bytes, "d2h_bytes_per_step": int(hits_step * 16),
                                         "note": "round 1's e2e arm: ASCII chunk up, 16-byte records per hit down, two host threads"},
            "probe_slow_seeds_per_step": acc_one["n_on_probe_sectors"] / steps,
            "offpath_mode": "index (walks from the starting loci materialised into the index)" if c0["offpath_mode"] == 2
                            else "walk (graph walked from the starting loci for every chunk)",
            "route": "fused one-pass kernel (seeding + probe + results)",
            "separate_kernels": {"note": "the same resident step (ASCII in, 32-byte records out) with set_option('fused', 0): "
                                         "per-kernel CUDA-event times and the seeds_on_paths probe kernel's own roofline",
                                 "ms_per_step": ms_sep / steps,
                                 "kernel_ms_per_step": {k_: acc_sep[k_] / steps for k_ in ("ms_pack", "ms_on", "ms_probe", "ms_resolve")},
                                 "roofline_probe": probe_roof},
            "clocks": clocks,
            "hits_per_read_histogram": {"bins": "reads with h hits in one step, h = 0..62, last bin >= 63; summed over ranks",
                                        "counts": [int(x) for x in hist]},
            "index": {"kmers": c0["n_index_kmers"], "entries": c0["n_index_entries"], "bytes": c0["index_bytes"],
                      "slot_bytes": c0["index_slot_bytes"], "build_ms": c0["ms_index_build"], "find_loci_ms": c0["ms_find_loci"],
                      "offpath_entries": c0["n_offpath_entries"], "offpath_walks": c0["n_offpath_walks"],
                      "stash_used": c0["index_stash_used"], "locus_codes": "by rank" if c0["code_by_rank"] else "by id"},
        }
        if drop_in:
            line["drop_in_psikt"] = drop_in
        if other:
            line["other_configs"] = other
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = reference_run(budget_s=8.0)
                line["cpu_baseline"] = {k_: r[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference",
                                        "sample": f"unavailable: {type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    W.close()
    if world > 1:
        dist.destroy_process_group()


def pcie_ceiling(torch, dev, up_bytes, down_bytes, reps=12):
    """What the host link of this GPU sustains for the e2e step's transfer sizes: pinned-memory copies alone and in both
    directions at once (two streams).  GB/s; the e2e step cannot beat max(up / h2d_bidir, down / d2h_bidir)."""
    hu = torch.empty(up_bytes, dtype=torch.uint8).pin_memory()
    hd = torch.empty(down_bytes, dtype=torch.uint8).pin_memory()
    du = torch.empty(up_bytes, dtype=torch.uint8, device=dev)
    dd = torch.empty(down_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed(do_up, do_down):
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        if do_up:
            with torch.cuda.stream(s1):
                ev[0].record()
                for _ in range(reps):
                    du.copy_(hu, non_blocking=True)
                ev[1].record()
        if do_down:
            with torch.cuda.stream(s2):
                ev[2].record()
                for _ in range(reps):
                    hd.copy_(dd, non_blocking=True)
                ev[3].record()
        torch.cuda.synchronize()
        up = up_bytes * reps / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9 if do_up else None
        down = down_bytes * reps / (ev[2].elapsed_time(ev[3]) * 1e-3) / 1e9 if do_down else None
        return up, down
    timed(True, True)
    h2d, _ = timed(True, False)
    _, d2h = timed(False, True)
    h2d_b, d2h_b = timed(True, True)
    return {"h2d_gbs": h2d, "d2h_gbs": d2h, "h2d_bidir_gbs": h2d_b, "d2h_bidir_gbs": d2h_b,
            "how": f"{reps} pinned-memory copies of {up_bytes} B up / {down_bytes} B down on two streams, CUDA events"}


def psikt_drop_in(W, n_reads=2_000_000, chunk=500_000):
    """The drop-in itself: psi_b200/bin/psikt from a reads file to the seed file (graph load and index build excluded:
    the CLI's own 'seed-finding' timer brackets the chunk loop: parse -> GPU -> write)."""
    import re
    import shutil
    from bench_support import synth
    psikt = ROOT / "psi_b200" / "bin" / "psikt"
    if not psikt.exists():
        return {"error": "psi_b200/bin/psikt not built"}
    td = tempfile.mkdtemp(prefix="psikt_bench_")
    try:
        gfa, fa, out, logf = (os.path.join(td, x) for x in ("g.gfa", "r.fa", "seeds.bin", "psikt.log"))
        W.g.write_gfa(gfa)
        rp, bases = synth.reads(W.g, n_reads, W.read_len, 4242)
        rec = np.empty((n_reads, 10 + W.read_len + 1), np.uint8)     # ">r0000000\n" + bases + "\n"
        rec[:, 0], rec[:, 1], rec[:, 9], rec[:, -1] = ord(">"), ord("r"), 10, 10
        ids = np.arange(n_reads)
        for c in range(7):
            rec[:, 8 - c] = 48 + (ids // 10 ** c) % 10
        rec[:, 10:-1] = bases.reshape(n_reads, W.read_len)
        rec.tofile(fa)
        t0 = time.time()
        p = subprocess.run([os.fspath(psikt), "-f", fa, "-l", str(W.k), "-n", str(N_PATHS), "-c", str(chunk), "-o", out, "-L", logf, "-q", gfa],
                           capture_output=True, text=True, timeout=900)
        wall = time.time() - t0
        if p.returncode != 0:
            return {"error": (p.stderr or "").strip()[-300:]}
        text = open(logf).read()
        m = re.search(r"Found seed in ([0-9.]+) s", text)
        t_find = float(m.group(1)) if m else None
        n_found = int(re.search(r"Total number of seeds found: (\d+)", text).group(1))
        return {"reads_per_s": n_reads / t_find if t_find else None, "seed_finding_s": t_find, "wall_s_with_graph_load_and_index": wall,
                "reads": n_reads, "chunk": chunk, "seeds_written": n_found, "output_bytes": os.path.getsize(out),
                "input_bytes": os.path.getsize(fa),
                "what": "psikt -f reads.fa -l K -n 16 -c CHUNK -o seeds.bin graph.gfa: FASTA parsing + 2-bit packing of the next chunk "
                        "on a second host thread while the first one waits for the GPU step, expands the 32-byte records and writes them"}
    finally:
        shutil.rmtree(td, ignore_errors=True)


def main_sharded(args):
    """BASELINE configs[2] / configs[4]: ONE read set sharded by read across the ranks (psi_b200/shard.py), graph and
    index replicated on every GPU, no data-path collective; NCCL only sums the counts.  Strong scaling: the read set is
    fixed (--reads-total), every rank takes a contiguous range holding the same number of bases and works through it
    in chunks of --reads.  Every rank checks its own results: completeness (reads are error-free walks of the graph, so
    every seed must hit) and soundness of a sample of records against the graph's labels."""
    import torch
    import torch.distributed as dist
    from bench_support import synth
    from psi_b200 import capi, shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    k, L, total, chunk = K, READ_LEN, args.reads_total, args.reads
    t0 = time.time()
    if world == 1:
        g = build_graph(args.shape)
        ps = g.pick_paths(N_PATHS, seed=1)
    else:
        # ONE host copy of the flattened graph and the picked paths per box: local rank 0 builds them and leaves the
        # arrays in /dev/shm, every rank maps them read-only (at whole-genome size a private copy per rank, ~50 GB,
        # would not fit the host's memory eight times)
        import shutil
        import types
        shm = Path("/dev/shm") / f"psi_b200_{args.shape}_{os.environ.get('MASTER_PORT', '0')}"
        names = ("seq_start", "seq", "row_ptr", "col", "internal_id", "coord_id", "path_ptr", "path_nodes", "path_head", "path_tail")
        if local == 0:
            g0 = build_graph(args.shape)
            ps0 = g0.pick_paths(N_PATHS, seed=1)
            shutil.rmtree(shm, ignore_errors=True)
            shm.mkdir(parents=True)
            for nm, arr in zip(names, (g0.seq_start, g0.seq, g0.row_ptr, g0.col, g0.internal_id, g0.coord_id, ps0.path_ptr, ps0.nodes,
                                       ps0.head_off, ps0.tail_trim)):
                np.save(shm / f"{nm}.npy", arr)
            np.save(shm / "meta.npy", np.array([g0.n_nodes, g0.n_edges, g0.n_bases, g0.n_paths], np.uint64))
            del g0, ps0
        dist.barrier()
        a = {nm: np.load(shm / f"{nm}.npy", mmap_mode="r") for nm in names}
        meta = np.load(shm / "meta.npy")
        g = types.SimpleNamespace(n_nodes=int(meta[0]), n_edges=int(meta[1]), n_bases=int(meta[2]), n_paths=int(meta[3]),
                                  seq_start=a["seq_start"], seq=a["seq"], row_ptr=a["row_ptr"], col=a["col"],
                                  internal_id=a["internal_id"], coord_id=a["coord_id"])
        ps = capi.PathSet(path_ptr=a["path_ptr"], nodes=a["path_nodes"], head_off=a["path_head"], tail_trim=a["path_tail"])
        dist.barrier()
        if local == 0:
            shutil.rmtree(shm, ignore_errors=True)      # the mappings keep the pages alive
    t_host = time.time() - t0
    ctx = capi.Context(k, local)
    ctx.set_option("offpath_mode", args.offpath_mode)
    for kv in args.opt:
        name, val = kv.split("=")
        ctx.set_option(name, int(val))
    ctx.set_graph(g, ids="internal")
    ctx.set_paths(ps)
    n_loci = ctx.find_loci()
    c0 = ctx.counters()
    free_b, total_b = torch.cuda.mem_get_info(dev)
    if rank == 0:
        log(f"[sharded] {args.shape}: graph {g.n_nodes} nodes / {g.n_bases} bp / {g.n_paths} components, host build {t_host:.0f} s; "
            f"index {c0['n_index_kmers']} k-mers at {c0['n_index_entries']} loci, {c0['index_bytes'] / 1e9:.2f} GB, slot "
            f"{c0['index_slot_bytes']} B, built in {c0['ms_index_build']:.0f} ms; {n_loci} starting loci, {c0['n_offpath_entries']} "
            f"off-path entries (mode {c0['offpath_mode']}) in {c0['ms_find_loci']:.0f} ms; device memory in use "
            f"{(total_b - free_b) / 1e9:.1f} of {total_b / 1e9:.0f} GB")
    # the ONE read set: blocks of a million reads, block b drawn with seed 7000 + b; a rank materialises only the blocks its
    # shard touches (shard_bounds cuts the whole set's offsets)
    read_ptr = np.arange(total + 1, dtype=np.uint64) * np.uint64(L)
    bounds = shard.shard_bounds(read_ptr, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    BLOCK = 1_000_000
    parts = []
    for b in range(lo // BLOCK, (hi + BLOCK - 1) // BLOCK):
        nb = min(BLOCK, total - b * BLOCK)
        _, bases = synth.reads(g, nb, L, 7000 + b)
        s, e = max(lo, b * BLOCK) - b * BLOCK, min(hi, b * BLOCK + nb) - b * BLOCK
        parts.append(bases[s * L:e * L])
    bases = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    n_mine = hi - lo
    per_read = (L - k) // k + 1
    chunks = []       # (first read id, n reads, pinned words)
    for s in range(0, n_mine, chunk):
        e = min(n_mine, s + chunk)
        pk = capi.Packed.pack(np.arange(e - s + 1, dtype=np.uint64) * np.uint64(L), bases[s * L:e * L], lo + s)
        chunks.append((lo + s, e - s, torch.from_numpy(pk.words.view(np.int64)).pin_memory()))
    n_pipes = 3
    pipes = [ctx] + [ctx.fork() for _ in range(n_pipes - 1)]
    d5_bits, d5 = ctx.dense5_layout()          # 5-byte results when the graph's loci fit 39 bits, else u32 id + u16 / u32 offset
    ob = 1 if d5 else ctx.dense_off_bytes()
    dense_h = [torch.empty(chunk * per_read * (4 + ob), dtype=torch.uint8).pin_memory() for _ in pipes]
    extra_h = [torch.empty((chunk * per_read // 8 + 4096, 4), dtype=torch.int32).pin_memory() for _ in pipes]
    DENSE = capi.ALL | (capi.DENSE5 if d5 else capi.DENSE)
    stats = {"hits": 0, "seeds_hit": 0, "checked": 0}
    rng = np.random.default_rng(5 + rank)

    def check(first, n, p):
        # completeness + soundness of a sample, straight from what arrived in pinned host memory
        ns, ne = pipes[p].dense_counts()
        assert ns == n * per_read, (ns, n, per_read)
        raw = dense_h[p].numpy()[:ns * (4 + ob)]
        dense = capi.dense5_planes(raw, ns, d5_bits) if d5 else capi.dense_planes(raw, ns, ob)
        hit = dense[:, 0] != capi.NIL32
        stats["seeds_hit"] += int(hit.sum())
        assert hit.all(), f"{int((~hit).sum())} seeds of error-free reads found nothing"
        if stats["checked"] < 4000:
            idx = rng.integers(0, ns, 500)
            ranks = np.searchsorted(g.internal_id, dense[idx, 0].astype(np.uint64))      # gum's internal ids grow with the rank
            assert np.array_equal(g.internal_id[ranks], dense[idx, 0].astype(np.uint64))
            for s_i, v in zip(idx, ranks):
                r, j = divmod(int(s_i), per_read)
                want = bases[((first - lo) + r) * L + j * k:((first - lo) + r) * L + j * k + k].tobytes()
                off = int(dense[s_i, 1] & 0x7FFFFFFF)
                assert want in _walks_from(g, int(v), off, k), "a record does not spell its seed in the graph"
            stats["checked"] += len(idx)

    def run(verify, resident):
        busy = [None] * n_pipes
        words_d = [w.to(dev) for _, _, w in chunks] if resident else None
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i, (first, n, w) in enumerate(chunks):
            p = i % n_pipes
            if busy[p] is not None:
                stats["hits"] += pipes[p].wait()
                if verify:
                    check(*busy[p], p)
            pipes[p].submit_chunk_packed_raw(n, n * L, L, (words_d[i] if resident else w).data_ptr(), first, k, on_device=resident)
            pipes[p].seeds_all_async(DENSE)
            if not resident:
                pipes[p].fetch_dense_async(dense_h[p].data_ptr(), chunk * per_read, extra_h[p].data_ptr(), extra_h[p].shape[0])
            busy[p] = (first, n)
        for p in range(n_pipes):
            if busy[p] is not None:
                stats["hits"] += pipes[p].wait()
                if verify:
                    check(*busy[p], p)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    run(True, False)                  # warm-up pass that also verifies every chunk
    hits_verified = stats["hits"]
    stats["hits"] = 0
    ms_e2e = run(False, False)
    hits_e2e = stats["hits"]
    stats["hits"] = 0
    ms_res = run(False, True) if n_mine * L / 4 < 0.25 * free_b else None
    counts, _ = shard.all_reduce_counts({"reads": n_mine, "hits": hits_e2e, "seeds": n_mine * per_read, "checked": stats["checked"]},
                                        np.zeros(shard.HIST_BINS, np.int64), device=dev)
    assert hits_e2e == hits_verified
    if rank == 0:
        assert counts["reads"] == total
        line = {"metric": f"reads/s (fully-sensitive seed finding, {args.shape}-shape graph, k={k})",
                "value": total / (ms_res * 1e-3) if ms_res else None, "unit": "reads/s", "n_gpus": world, "steps": len(chunks), "warmup": len(chunks),
                "ms_per_step": (ms_res or ms_e2e) / max(len(chunks), 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u8/u64", "data": "synthetic",
                "config": {"workload": f"synthetic {args.shape}-shape graph ({g.n_bases} bp, {g.n_nodes} nodes, {g.n_paths} components, "
                                       f"{N_PATHS} paths each, {n_loci} starting loci), ONE set of {total} x {L} bp reads sharded by read over "
                                       f"{world} GPU(s) (psi_b200/shard.py), chunks of {chunk} reads, k={k}, d={k}",
                           "sharding": "contiguous read ranges of equal base count per rank, graph + index replicated, NCCL all-reduce of counts only"},
                "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "reads/s", "ms_total": ms_e2e,
                        "h2d_bytes_per_step": int(chunks[0][2].numel() * 8) if chunks else 0,
                        "d2h_bytes_per_step": int(chunk * per_read * (4 + ob)),
                        "result_format": "PSI_B200_DENSE5 (5 bytes per seed)" if d5 else f"PSI_B200_DENSE ({4 + ob} bytes per seed)"},
                "seeds_per_s": counts["hits"] / ((ms_res or ms_e2e) * 1e-3), "hits": counts["hits"], "query_seeds": counts["seeds"],
                "verified": f"every seed of every (error-free) read hit; {counts['checked']} sampled records spell their seed in the graph",
                "index": {"kmers": c0["n_index_kmers"], "entries": c0["n_index_entries"], "bytes": c0["index_bytes"],
                          "slot_bytes": c0["index_slot_bytes"], "build_ms": c0["ms_index_build"], "find_loci_ms": c0["ms_find_loci"],
                          "offpath_entries": c0["n_offpath_entries"], "device_bytes_in_use": int(total_b - free_b),
                          "build_slices": c0["index_build_slices"], "host_graph_and_paths_s": t_host}}
        print(json.dumps(line), flush=True)
    for cx in pipes[1:]:
        cx.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def _walks_from(g, v, off, k):
    """All k-mers spelt by forward walks from (node rank v, offset off) -- the soundness check of main_sharded."""
    out, stack = [], [(v, off, b"")]
    while stack:
        v, o, acc = stack.pop()
        acc += g.seq[int(g.seq_start[v]) + o:int(g.seq_start[v + 1])].tobytes()[:k - len(acc)]
        if len(acc) == k:
            out.append(acc)
            continue
        for e in range(int(g.row_ptr[v]), int(g.row_ptr[v + 1])):
            stack.append((int(g.col[e]), 0, acc))
    return out


def other_configs(torch, dev, run_async, run_sync, capi):
    """The other single-GPU shapes BASELINE.json names, one short measurement each (resident 2-bit chunks -> dense
    results, 4 chunks in flight): configs[2]'s read shape (150 bp) on the chr22-shape graph, configs[3] (MHC-like
    region, k = 32) in index mode and in walk mode (the reference's scheme: graph walked per chunk)."""
    out = {}
    DENSE = capi.ALL | capi.DENSE
    k0, len0 = K, READ_LEN
    for name, shape, k, read_len, n_reads, mode in (("chr22_150bp", "chr22", 20, 150, 1_000_000, 0),
                                                     ("mhc_k32", "mhc", 32, 150, 1_000_000, 0),
                                                     ("mhc_k32_walk_mode", "mhc", 32, 150, 200_000, 1)):
        try:
            globals().update(K=k, READ_LEN=read_len)
            W = Workload(torch, dev, 0, shape, k, read_len, n_reads, 2, N_PATHS, 4, mode)
            if mode == 0:
                d5 = W.ctx.dense5_layout()[1]
                fast = capi.ALL | (capi.DENSE5 if d5 else capi.DENSE)
                ms, hits, launches, kms = run_async(W, 6, 3, 4, "packed", "device", fast, False)
                ms_e, hits_e, _, _ = run_async(W, 6, 3, 4, "packed", "host", fast, True)
                out[name] = {"result_bytes_per_seed": 5 if d5 else 4 + W.ctx.dense_off_bytes(), "value": n_reads * 6 / (ms * 1e-3), "unit": "reads/s", "ms_per_step": ms / 6, "fused_kernel_ms": kms,
                             "query_seeds_per_s": W.n_seeds * 6 / (ms * 1e-3), "hits_per_step": hits / 6,
                             "e2e": n_reads * 6 / (ms_e * 1e-3), "index_bytes": W.c0["index_bytes"], "slot_bytes": W.c0["index_slot_bytes"],
                             "starting_loci": W.n_loci, "offpath_entries": W.c0["n_offpath_entries"],
                             "workload": f"{shape}-shape graph, {n_reads} x {read_len} bp reads per step, k={k}"}
            else:
                ms, hits, acc, launches = run_sync(W, 4, 3, "packed", capi.ALL)
                out[name] = {"value": n_reads * 4 / (ms * 1e-3), "unit": "reads/s", "ms_per_step": ms / 4,
                             "kernel_ms_per_step": {k_: acc[k_] / 4 for k_ in ("ms_pack", "ms_read_index", "ms_on", "ms_off", "ms_resolve")},
                             "walks_per_step": acc["n_walks"] / 4, "starting_loci": W.n_loci,
                             "workload": f"{shape}-shape graph walked from its starting loci for every chunk of {n_reads} x {read_len} bp reads, k={k}"}
            if name == "chr22_150bp":
                out["distance_index_300_500"] = distance_bench(torch, dev, capi, W.g, 300, 500)
                out["mem_mode_chr22"] = mem_bench(torch, dev, capi, W)
            W.close()
        except Exception as e:   # an extra line must not take the headline down
            out[name] = {"error": f"{type(e).__name__}: {e}"}
    globals().update(K=k0, READ_LEN=len0)
    return out


def mem_bench(torch, dev, capi, W, n_reads=200_000):
    """MEM mode (SeedFinder::seeds_on_paths(sequence, cb) -> find_mems) on the same graph and paths: the
    suffix table of the path text is built, then one chunk of reads is scanned (hits stay on the device)."""
    try:
        import ctypes as C
        ctx = capi.Context(W.k, dev.index)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx.set_graph(W.g, ids="internal")
        t0 = time.perf_counter()
        ctx.build_mem_index(W.ps)
        build_s = time.perf_counter() - t0
        L = W.read_len
        words = W.words_d[0]
        n_hits = C.c_uint64()
        times = []
        for _ in range(3):
            ctx.submit_chunk_packed_raw(n_reads, n_reads * L, L, words.data_ptr(), 0, W.k, on_device=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx._ck(capi.lib().psi_b200_find_mems(ctx._h, 0, C.byref(n_hits)))
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = sorted(times)[1]
        res = {"reads_per_s": n_reads / (ms * 1e-3), "ms_per_chunk": ms, "reads": n_reads, "read_len": L, "min_len": W.k, "mems": n_hits.value,
               "index_build_s": build_s,
               "workload": f"chr22-shape graph, {len(W.ps.path_ptr) - 1} indexed paths, {n_reads} x {L} bp error-free reads, minimum MEM length {W.k}"}
        ctx.close()
        return res
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"}


def distance_bench(torch, dev, capi, g, dmin, dmax, n_pairs=4_000_000):
    """Paired-end distance verification (SeedFinder::create_distance_index / verify_distance) on the same
    graph: index build, then n_pairs locus pairs resident in HBM answered against the materialised rows and by
    enumeration; pairs are drawn so that roughly a third fall inside the window."""
    try:
        rng = np.random.default_rng(5)
        start = np.asarray(g.seq_start, np.int64)
        a = rng.integers(0, int(start[-1]) - 2 * dmax - 64, n_pairs)
        b = a + rng.integers(0, int(1.5 * dmax), n_pairs)
        rv, ru = np.searchsorted(start, a, side="right") - 1, np.searchsorted(start, b, side="right") - 1
        pairs = np.stack([rv, a - start[rv], ru, b - start[ru]], axis=1).astype(np.uint32)
        d_pairs = torch.from_numpy(pairs.view(np.int32)).to(dev)
        d_out = torch.zeros(n_pairs, dtype=torch.uint8, device=dev)
        res = {"workload": f"chr22-shape graph, window {dmin}..{dmax}, {n_pairs} locus pairs resident in HBM"}
        for mode, key in ((2, "rows"), (1, "enumerate")):
            ctx = capi.Context(K, dev.index)
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            ctx.set_graph(g, ids="internal")
            ctx.set_option("dindex_mode", mode)
            t0 = time.perf_counter()
            ctx.create_distance_index(dmin, dmax)
            build_s = time.perf_counter() - t0
            c = ctx.counters()
            n_q = n_pairs if mode == 2 else n_pairs // 8
            ctx.verify_distance_device(n_q, d_pairs.data_ptr(), d_out.data_ptr())      # warm-up
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ctx.verify_distance_device(n_q, d_pairs.data_ptr(), d_out.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            res[key] = {"queries_per_s": n_q / (ms * 1e-3), "ms_per_call": ms, "queries_per_call": n_q,
                        "inside_window": int(d_out[:n_q].sum().item())}
            if mode == 2:
                res["index"] = {"entries": c["n_dindex_entries"], "bytes": c["dindex_bytes"], "build_ms_device": c["ms_dindex_build"],
                                "build_s_wall": build_s, "bytes_per_query": 16 + 1 + 12 + 8 * c["n_dindex_entries"] / max(c["n_nodes"], 1)}
            ctx.close()
        return res
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="psi_b200", choices=["psi_b200", "reference"])
    ap.add_argument("--shape", default="chr22", help="bench_support.synth.SHAPES key")
    ap.add_argument("--reads", type=int, default=READS_PER_BATCH)
    ap.add_argument("--k", type=int, default=K, help="seed length (other BASELINE configs: 32 with --shape mhc)")
    ap.add_argument("--read-len", type=int, default=READ_LEN, help="read length (150 for BASELINE configs[2..4])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipelines", type=int, default=6, help="contexts (forks sharing one index) kept in flight by the one host thread")
    ap.add_argument("--reads-total", type=int, default=0,
                    help="> 0: strong scaling -- ONE read set of this many reads sharded over the ranks in chunks of --reads "
                         "(BASELINE configs[2]: --reads-total 10000000 --read-len 150; configs[4]: --shape wg_1_4 ...)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short runs of the other BASELINE configs (N = 1 only)")
    ap.add_argument("--opt", action="append", default=[], help="name=value passed to psi_b200_set_option (tuning experiments)")
    ap.add_argument("--offpath-mode", type=int, default=0, help="0 auto, 1 walk per chunk, 2 materialise (psi_b200_set_option)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "psi_b200" else args.warmup
    globals().update(K=args.k, READ_LEN=args.read_len)
    if args.impl == "reference":
        main_reference(args)
    elif args.reads_total > 0:
        main_sharded(args)
    else:
        main_gpu(args)


if __name__ == "__main__":
    main()
