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
    "seeds_fused_kernel": (897.6e6, "profiles/r01zn_fused_ncu_raw.csv: seeds_fused_kernel<8, 5, 4>, dram__bytes_read.sum 780.6 MB + "
                                    "dram__bytes_write.sum 117.0 MB (mean of 2 launches)"),
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
            self.rows.append(line.strip())

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
        for r in self.rows:
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
    a = synth.graph_arrays(**synth.SHAPES[shape])
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
        # the reference picks its paths with std::random_device, so a crash need not repeat: one more try at the same
        # process count, then halve
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

def main_gpu(args):
    import torch
    import torch.distributed as dist
    from bench_support import synth
    from psi_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    t0 = time.time()
    g = build_graph(args.shape)
    ps = g.pick_paths(N_PATHS, seed=1)
    ctx = capi.Context(K, local)
    ctx.set_option("offpath_mode", args.offpath_mode)
    for kv in args.opt:
        name, val = kv.split("=")
        ctx.set_option(name, int(val))
    if args.blocking_sync:
        ctx.set_option("blocking_sync", 1)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_graph(g, ids="internal")
    ctx.set_paths(ps)
    n_loci = ctx.find_loci()
    c0 = ctx.counters()
    if rank == 0:
        log(f"[bench] graph {g.n_nodes} nodes / {g.n_bases} bp; index {c0['n_index_kmers']} k-mers, "
            f"{c0['index_bytes'] / 1e6:.0f} MB, slot {c0['index_slot_bytes']} B, build {c0['ms_index_build']:.0f} ms; "
            f"{n_loci} starting loci, {c0['n_offpath_walks']} uncovered walks -> {c0['n_offpath_entries']} off-path entries "
            f"(mode {c0['offpath_mode']}) in {c0['ms_find_loci']:.0f} ms; setup {time.time() - t0:.1f} s")

    # distinct read batches per rank (weak scaling: every GPU processes its own shard of the read set)
    n_reads = args.reads
    batches_h, batches_d = [], []
    for b in range(N_BATCHES):
        rp, bases = synth.reads(g, n_reads, READ_LEN, 1002 + 1000 * rank + b)
        hp = torch.from_numpy(rp.view(np.int64)).pin_memory()
        hb = torch.from_numpy(bases).pin_memory()
        batches_h.append((hp, hb))
        batches_d.append((hp.to(dev), hb.to(dev)))
    # e2e runs two pipelines (the context and a fork sharing its resident index) from two host threads, so that
    # the upload of one batch overlaps the kernels and the download of the previous one (PCIe is full duplex)
    n_pipes = max(1, args.pipelines)                 # e2e: two pipelines keep both PCIe directions busy; more only contend
    # resident inputs: four pipelines let one chunk's kernel tail overlap the next one's head (5.4 vs 4.7 G reads/s on one
    # GPU), but every pipeline is a host thread that spins in its stream synchronisation: with fewer than 8 host cores
    # per rank (8 ranks on this 32-core box) four of them get in each other's way (35.2 vs 37.1 G reads/s at 8 GPUs)
    n_vpipes = args.value_pipelines if args.value_pipelines > 0 else (4 if (os.cpu_count() or 1) // world >= 8 else 2)
    pipes = [ctx] + [ctx.fork() for _ in range(max(n_pipes, n_vpipes) - 1)]
    rec_hosts = [torch.empty((8 * n_reads, 4), dtype=torch.int64).pin_memory() for _ in range(n_pipes)]   # room for the seed records
    torch.cuda.synchronize()

    def step_device(i):
        dp, db = batches_d[i % N_BATCHES]
        ctx.submit_chunk_device(n_reads, dp.data_ptr(), db.data_ptr(), db.numel(), rank * n_reads, K)
        return ctx.seeds_all(capi.ALL)

    def step_e2e(i, p=0, compact=True):
        hp, hb = batches_h[i % N_BATCHES]
        cx, rec_host = pipes[p], rec_hosts[p]
        cx.submit_chunk_ptr(n_reads, hp.data_ptr(), hb.data_ptr(), rank * n_reads, K)
        if compact:      # 4 x u32 records: the same fields, half the bytes over PCIe
            cx.seeds_all(capi.ALL | capi.COMPACT)
            return cx.fetch32_into(rec_host.data_ptr(), rec_host.shape[0])
        cx.seeds_all(capi.ALL)
        return cx.fetch_into(rec_host.data_ptr(), rec_host.shape[0])

    def step_resident(i, p=0):
        dp, db = batches_d[i % N_BATCHES]
        pipes[p].submit_chunk_device(n_reads, dp.data_ptr(), db.data_ptr(), db.numel(), rank * n_reads, K)
        return pipes[p].seeds_all(capi.ALL)

    def timed_e2e(steps, warmup, step_fn=step_e2e, n_pipes=n_pipes):
        """K steps through the C-ABI, round-robin over the first n_pipes pipelines, one host thread each."""
        for i in range(max(warmup, n_pipes)):
            step_fn(i, i % n_pipes)
        barrier()
        for cx in pipes:
            cx.reset_counters()
        hits = [0] * n_pipes
        errors = []

        def work(p):
            try:
                for i in range(p, steps, n_pipes):
                    hits[p] += step_fn(warmup + i, p)
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
        ms = e0.elapsed_time(e1)
        launches = sum(cx.counters()["launches"] for cx in pipes)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, sum(hits), launches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup):
        for i in range(warmup):
            step_fn(i)
        barrier()
        ctx.reset_counters()
        acc = {"ms_on": 0.0, "ms_probe": 0.0, "ms_off": 0.0, "ms_pack": 0.0, "ms_read_index": 0.0, "ms_resolve": 0.0, "ms_h2d": 0.0,
               "ms_d2h": 0.0, "n_hits_on": 0, "n_hits": 0, "n_seeds": 0, "n_walks": 0, "n_on_probe_sectors": 0}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hits = 0
        for i in range(steps):
            hits += step_fn(warmup + i)
            c = ctx.counters()          # syncs the stream; per-kernel CUDA-event times of this step
            for k_ in acc:
                acc[k_] += c[k_]
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ctx.counters()["launches"]
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, hits, acc, launches

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, hits_dev, acc, launches = timed(step_device, args.steps, args.warmup)
    c_last = ctx.counters()
    ms_e2e, hits_e2e, launches_e2e = timed_e2e(args.steps, args.warmup)
    ms_e2e_wide, hits_e2e_wide, _ = timed_e2e(args.steps, args.warmup, lambda i, p=0: step_e2e(i, p, compact=False))
    # the resident-input step again with the pipelines running concurrently (kernels of one chunk fill the launch and
    # latency gaps of the other); `value` stays the single-pipeline figure the per-kernel timers belong to
    ms_pipe, hits_pipe, launches_pipe = timed_e2e(args.steps, args.warmup, step_resident, n_vpipes)
    assert hits_e2e == hits_e2e_wide == hits_pipe == hits_dev, (hits_e2e, hits_e2e_wide, hits_pipe, hits_dev)
    # the same resident step through the separate seeding / probe / resolve kernels: per-kernel times, and the
    # seeds_on_paths probe alone for its own roofline
    fused_run = bool(c_last["fused"])
    if fused_run:
        ctx.set_option("fused", 0)
        ms_sep, hits_sep, acc_sep, _ = timed(step_device, args.steps, args.warmup)
        ctx.set_option("fused", 1)
        assert hits_sep == hits_dev, (hits_sep, hits_dev)
    else:
        ms_sep, acc_sep = ms_dev, acc
    clocks = sampler.stop() if rank == 0 else None

    # per-shard counts and a hits-per-read histogram, reduced with NCCL (the only collective on this path)
    from psi_b200 import shard

    class _DevArray:   # view the device records of the last step as a torch tensor (no copy)
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n, 4), "typestr": "<i8", "data": (ptr, False), "version": 2}
    step_device(0)
    ptr, n_rec = ctx.fetch_device()
    if n_rec:
        rec = torch.as_tensor(_DevArray(ptr, n_rec), device=dev)
        per_read = torch.bincount(rec[:, 2] - rank * n_reads, minlength=n_reads)[:n_reads]
        hist = torch.bincount(torch.clamp(per_read, max=shard.HIST_BINS - 1), minlength=shard.HIST_BINS).cpu().numpy()
    else:
        hist = np.zeros(shard.HIST_BINS, np.int64)
        hist[0] = n_reads
    counts, hist = shard.all_reduce_counts({"reads": n_reads * args.steps, "seeds": acc["n_seeds"], "hits": hits_dev,
                                            "hits_on": acc["n_hits_on"], "walks": acc["n_walks"]}, hist, device=dev)
    tot = [counts["reads"], counts["seeds"], counts["hits"], counts["hits_on"], counts["walks"]]

    if rank == 0:
        peak, peak_src = peaks()
        reads_total, seeds_total, hits_total, hits_on_total, walks_total = tot
        # value: K steps issued round-robin over --value-pipelines contexts (the context and its forks share one resident index; a
        # pipeline's host-side launch and read-back gaps are filled by the other's kernels) -- the way the library is
        # meant to be driven, and the way e2e is measured.  The one-pipeline loop, whose per-kernel CUDA-event times
        # feed `kernel_ms_per_step` and `roofline`, is reported beside it.
        ms_value = ms_pipe if n_vpipes > 1 else ms_dev
        value = reads_total / (ms_value * 1e-3)
        # roofline of the dominant kernel (seeds_on_paths probe), rank 0's launches.  Algorithmic bytes per launch =
        # seeds x (8 B packed k-mer + 128 B = ONE index bucket line, the DRAM access unit: profiles/r01c_gather_peak.md)
        # + hits x 8 B compact record.  The 32-B-sector accounting of SURVEY 8d (40 B per seed) is reported beside it.
        n_probe_hits = acc_sep["n_hits"] if c_last["offpath_mode"] == 2 else acc_sep["n_hits_on"]
        alg_bytes = (acc_sep["n_seeds"] * 136 + n_probe_hits * 8) / args.steps
        alg_bytes_sector = (acc_sep["n_seeds"] * 40 + n_probe_hits * 8) / args.steps
        on_ms = acc_sep["ms_probe"] / args.steps  # the probe kernel alone; ms_on = probe + slow-queue kernel
        achieved = alg_bytes / (on_ms * 1e-3) / 1e9 if on_ms > 0 else 0.0
        achieved_sector = alg_bytes_sector / (on_ms * 1e-3) / 1e9 if on_ms > 0 else 0.0
        per_step = {k_: acc[k_] / args.steps for k_ in ("ms_pack", "ms_on", "ms_probe", "ms_read_index", "ms_off", "ms_resolve")}
        probe_roof = {"bound": "hbm", "kernel": "seeds_on_paths_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                      "frac": achieved / peak, "traffic": TRAFFIC["seeds_on_paths_kernel"][0], "peak_source": peak_src,
                      "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": on_ms,
                      "bytes_per_seed": "8 B k-mer + 128 B bucket line + 8 B per hit",
                      "sector_accounting": {"bytes_per_seed": "8 B k-mer + 32 B sector + 8 B per hit (SURVEY 8d, P=1)",
                                            "achieved": achieved_sector, "frac": achieved_sector / peak},
                      "probes_per_s": acc_sep["n_seeds"] / args.steps / (on_ms * 1e-3) if on_ms > 0 else 0.0,
                      "random_line_ceiling_probes_per_s": 4.0e10,
                      "traffic_source": TRAFFIC["seeds_on_paths_kernel"][1]}
        if fused_run:
            # the fused kernel is the step: chunk bytes + read offsets in, one bucket line per seed, one 32-byte record
            # + 1 kind byte per hit out (the position -> node gathers hit L2-resident arrays and are not counted)
            hb0 = batches_h[0][1]
            f_bytes = hb0.numel() + 8 * (n_reads + 1) + (acc["n_seeds"] * 128 + acc["n_hits"] * 33) / args.steps
            f_sector = hb0.numel() + 8 * (n_reads + 1) + (acc["n_seeds"] * 32 + acc["n_hits"] * 33) / args.steps
            f_ms = acc["ms_probe"] / args.steps
            f_ach = f_bytes / (f_ms * 1e-3) / 1e9 if f_ms > 0 else 0.0
            roof = {"bound": "hbm", "kernel": "seeds_fused_kernel", "achieved": f_ach, "peak": peak, "unit": "GB/s",
                    "frac": f_ach / peak, "traffic": TRAFFIC["seeds_fused_kernel"][0], "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": f_bytes, "launch_ms": f_ms,
                    "bytes_per_read": "read bytes + 8 B offset + per seed one 128 B bucket line + per hit a 32 B record and 1 kind byte",
                    "sector_accounting": {"bytes": "the same with 32 B per seed instead of the 128 B line (SURVEY 8d, P=1)",
                                          "achieved": f_sector / (f_ms * 1e-3) / 1e9 if f_ms > 0 else 0.0,
                                          "frac": f_sector / (f_ms * 1e-3) / 1e9 / peak if f_ms > 0 else 0.0},
                    "probes_per_s": acc["n_seeds"] / args.steps / (f_ms * 1e-3) if f_ms > 0 else 0.0,
                    "random_line_ceiling_probes_per_s": 4.0e10,
                    "traffic_source": TRAFFIC["seeds_fused_kernel"][1]}
        else:
            roof = probe_roof
        e2e_value = n_reads * args.steps * world / (ms_e2e * 1e-3)
        hp, hb = batches_h[0]
        line = {
            "metric": f"reads/s (fully-sensitive seed finding, {args.shape}-shape graph, k={K})",
            "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/u64", "data": "synthetic",
            "config": {"workload": f"synthetic {args.shape}-shape graph ({g.n_bases} bp, {g.n_nodes} nodes, {N_PATHS} paths, "
                                   f"{n_loci} starting loci), {n_reads} x {READ_LEN} bp reads per GPU per step, k={K}, d={K}",
                       "l2": f"{N_BATCHES} distinct read batches cycled ({N_BATCHES * n_reads * READ_LEN / 1e6:.0f} MB) and a "
                             f"{c0['index_bytes'] / 1e6:.0f} MB index: inputs larger than L2",
                       "sharding": "reads sharded by rank, graph + index replicated, NCCL all-reduce of counts only",
                       "pipelines": f"contexts per GPU sharing one resident index, steps issued round-robin: {n_vpipes} for value "
                                    f"(inputs resident), {n_pipes} for e2e (more only contend for PCIe)"},
            "seeds_per_s": hits_total / (ms_value * 1e-3), "query_seeds_per_s": seeds_total / (ms_value * 1e-3),
            "kernel_ms_per_step": per_step, "probe_slow_seeds_per_step": acc["n_on_probe_sectors"] / args.steps,
            "e2e": {"value": e2e_value, "unit": "reads/s",
                    "h2d_bytes_per_step": int(hb.numel() + hp.numel() * 8),
                    "d2h_bytes_per_step": int(hits_e2e / args.steps * 16), "ms_per_step": ms_e2e / args.steps,
                    "pipelines": n_pipes,
                    "records": "4 x u32 {node_id, node_offset, read_id, read_offset} per hit (PSI_B200_COMPACT + psi_b200_fetch32)"},
            "e2e_wide_records": {"value": n_reads * args.steps * world / (ms_e2e_wide * 1e-3), "unit": "reads/s",
                                 "h2d_bytes_per_step": int(hb.numel() + hp.numel() * 8),
                                 "d2h_bytes_per_step": int(hits_e2e_wide / args.steps * 32),
                                 "ms_per_step": ms_e2e_wide / args.steps, "pipelines": n_pipes,
                                 "records": "4 x u64 per hit, the byte layout psikt writes (psi_b200_fetch)"},
            "value_one_pipeline": {"value": reads_total / (ms_dev * 1e-3), "unit": "reads/s", "ms_per_step": ms_dev / args.steps,
                                   "note": "the same K steps on ONE context, each step synchronised before the next is issued; "
                                           "kernel_ms_per_step and roofline are this loop's CUDA-event times"},
            "gpu_launches": int(launches_pipe if n_vpipes > 1 else launches), "gpu_launches_e2e": int(launches_e2e),
            "offpath_mode": "index (walks from the starting loci materialised into the index)" if c_last["offpath_mode"] == 2
                            else "walk (graph walked from the starting loci for every chunk)",
            "roofline": roof,
            "route": "fused one-pass kernel (seeding + probe + records)" if fused_run else "separate seeding / probe / resolve kernels",
            "separate_kernels": {"note": "the same resident step with set_option('fused', 0): per-kernel CUDA-event times and the "
                                         "seeds_on_paths probe kernel's own roofline",
                                 "ms_per_step": ms_sep / args.steps,
                                 "kernel_ms_per_step": {k_: acc_sep[k_] / args.steps for k_ in ("ms_pack", "ms_on", "ms_probe", "ms_resolve")},
                                 "roofline_probe": probe_roof},
            "clocks": clocks,
            "hits_per_read_histogram": {"bins": "reads with h hits in one step, h = 0..62, last bin >= 63; summed over ranks",
                                        "counts": [int(x) for x in hist]},
            "index": {"kmers": c0["n_index_kmers"], "entries": c0["n_index_entries"], "bytes": c0["index_bytes"],
                      "slot_bytes": c0["index_slot_bytes"], "build_ms": c0["ms_index_build"], "find_loci_ms": c0["ms_find_loci"],
                      "offpath_entries": c0["n_offpath_entries"], "offpath_walks": c0["n_offpath_walks"],
                      "stash_used": c0["index_stash_used"]},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                r = reference_run(budget_s=8.0)
                line["cpu_baseline"] = {k_: r[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference",
                                        "sample": f"unavailable: {type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    for cx in pipes[1:]:
        cx.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--pipelines", type=int, default=2, help="e2e: contexts (forks sharing one index) driven concurrently")
    ap.add_argument("--value-pipelines", type=int, default=0,
                    help="value (inputs resident in HBM): contexts driven concurrently; 0 = 4 with >= 8 host cores per rank, else 2")
    ap.add_argument("--blocking-sync", type=int, default=0, help="1: pipelines sleep on a blocking event instead of spinning")
    ap.add_argument("--opt", action="append", default=[], help="name=value passed to psi_b200_set_option (tuning experiments)")
    ap.add_argument("--offpath-mode", type=int, default=0, help="0 auto, 1 walk per chunk, 2 materialise (psi_b200_set_option)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "psi_b200" else args.warmup
    globals().update(K=args.k, READ_LEN=args.read_len)
    if args.impl == "reference":
        main_reference(args)
    else:
        main_gpu(args)


if __name__ == "__main__":
    main()
