#!/usr/bin/env python
"""bench.py -- all-pairs anchor-distance throughput (BASELINE.json metric) on N B200s.

A *step* is one pass of the hot path over one batch: `rows` subjects of the all-pairs matrix
(index construction for each subject + its anchor walk against every genome of the pool).
Workload (default): BASELINE.json configs[3] -- 3085 synthetic 2.1 Mbp genomes, star phylogeny,
d_k ~ U[0.005, 0.02] from the base, JC. The whole pool is resident on every GPU (it is the
"weights" of the job); subjects are sharded over ranks (weak scaling: every rank does `rows`
subjects per step) and finished row blocks are gathered to rank 0 with NCCL.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU code on the host cores

value   pairs/s with the pool already packed in HBM, device-timed (CUDA events, max over ranks)
e2e     pairs/s through the host-buffer C ABI: every step uploads the pool from pinned host
        memory (andi_pool_set_host), computes the rows and reads them back to the host
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "all-pairs genome-pairs/sec at 3085x2.1Mbp"
UNIT = "pairs/s"

WORKLOADS = {
    # name: (genomes, length, d_lo, d_hi, base seed, model)  -- SURVEY.md 8d
    "c4": (3085, 2_100_000, 0.005, 0.02, 3085, "JC"),
    "c2": (29, 5_000_000, 0.01, 0.05, 29, "JC"),
    "c1": (2, 100_000, 0.0099, 0.0099, 1729, "JC"),
    "c3": (109, 5_000_000, 0.01, 0.05, 109, "KIMURA"),
    "c5": (16, 120_000_000, 0.01, 0.05, 16, "JC"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--genomes", type=int, default=0, help="override the number of genomes (experiments only)")
    ap.add_argument("--length", type=int, default=0, help="override the genome length (experiments only)")
    ap.add_argument("--contigs", type=int, default=0,
                    help="join-mode shape: every genome is this many contigs joined by '!' (experiments only)")
    ap.add_argument("--repeats", type=int, default=0,
                    help="realistic repeats: this many copies of a 1.5 kbp element and 7 copies of a 5 kbp operon in the base "
                         "genome (exact repeats beyond the direct sort / LCP caps: prefix doubling + phi LCP; experiments only)")
    ap.add_argument("--rows", type=int, default=0, help="subjects per step and rank (default: one full walk batch)")
    ap.add_argument("--model", default="")
    ap.add_argument("--divergence", default="", help="lo,hi: distance of every genome from the base, uniform (experiments only; "
                                                     "1e-5,1e-4 = outbreak isolates: anchors of tens of kilobases)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-full", action="store_true", help="skip the whole-matrix (strong scaling) leg")
    ap.add_argument("--cpu-queries", type=int, default=192, help="queries per subject in the bounded CPU sample")
    return ap.parse_args()


def workload_of(args):
    g, ln, lo, hi, seed, model = WORKLOADS[args.workload]
    if args.genomes:
        g = args.genomes
    if args.length:
        ln = args.length
    if args.model:
        model = args.model
    if args.divergence:
        lo, hi = (float(x) for x in args.divergence.split(","))
    return g, ln, lo, hi, seed, model


def divergences(g, lo, hi, seed):
    return np.random.default_rng(seed ^ 0x5EED).uniform(lo, hi, size=g)


# ----------------------------------------------------------------------------- synthetic pool

def make_pool_device(g, ln, lo, hi, seed, device, contigs=0, repeats=0):
    """Star phylogeny on the device (shape of test/test_fasta.cxx): uniform base genome, genome k
    gets exactly round(len * d_k) substitutions at distinct uniform positions."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    stride = (ln + 15) // 16 * 16
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    base = torch.randint(0, 4, (ln,), dtype=torch.uint8, device=device, generator=gen)
    if repeats:  # IS-element-like and rRNA-operon-like exact repeats (bacterial genomes carry both)
        element, operon = base[:1500].clone(), base[2000:7000].clone()
        for at in torch.randint(10000, ln - 10000, (repeats,), device=device, generator=gen).tolist():
            base[at : at + 1500] = element
        for at in torch.randint(10000, ln - 10000, (7,), device=device, generator=gen).tolist():
            base[at : at + 5000] = operon
    chars = torch.zeros(g * stride, dtype=torch.uint8, device=device)
    d = divergences(g, lo, hi, seed)
    for k in range(g):
        nmut = int(round(ln * float(d[k])))
        codes = base.clone()
        if nmut:
            pos = torch.randperm(ln, device=device, generator=gen)[:nmut]
            shift = torch.randint(1, 4, (nmut,), dtype=torch.uint8, device=device, generator=gen)
            codes[pos] = (codes[pos] + shift) & 3
        chars[k * stride : k * stride + ln] = lut[codes.long()]
        if contigs > 1:  # '!' between contigs, as join mode writes them (src/io.c / sequence.c)
            cut = torch.randint(1, ln - 1, (contigs - 1,), device=device, generator=gen)
            chars[k * stride + cut] = ord("!")
    offsets = [k * stride for k in range(g)]
    return chars, offsets, [ln] * g, d


# ----------------------------------------------------------------------------- clocks

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in Path(self.f.name).read_text().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
            except ValueError:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU reference arm

def cpu_sample(host_seq, n_total, subjects, q_sample, model, threads):
    """Time the reference's own CPU code (oracle/_ref; else the single-core oracle port) on a
    bounded sample: `subjects` (one per thread) x q_sample queries each, FAST mode
    (src/dist_hack.h:8). Returns pairs/s extrapolated to full rows of n_total-1 queries:
        threads / (esa_seconds_per_subject / (n_total-1) + walk_seconds_per_pair)."""
    import oracle

    rng = np.random.default_rng(1234 + subjects[0])
    qids = [int(q) for q in rng.choice(n_total, size=min(q_sample, n_total - 1), replace=False) if q not in subjects]
    seqs = [host_seq(i) for i in subjects] + [host_seq(j) for j in qids]
    S = len(subjects)
    ref_cells = None
    if oracle.ref_available():
        ref_cells, t = oracle.ref_rows(seqs, model, s_begin=0, s_end=S, threads=threads)
        kind = "reference"
        esa_per_subject = t["esa_s"] / S
        walk_per_pair = t["walk_s"] / (S * (len(seqs) - 1))
        wall = t["wall_s"]
    else:
        t0 = time.perf_counter()
        ref_cells = oracle.rows(seqs, model, s_begin=0, s_end=1)
        wall = time.perf_counter() - t0
        kind, threads, S = "port", 1, 1
        esa_per_subject, walk_per_pair = 0.0, wall / (len(seqs) - 1)
    value = threads / (esa_per_subject / max(1, n_total - 1) + walk_per_pair)
    return {
        "value": value, "unit": UNIT, "cores": threads, "kind": kind,
        "sample": f"{S} subjects x {len(seqs) - 1} queries of the same pool, {wall:.1f} s wall; "
                  f"esa_init {esa_per_subject:.3f} s/subject (SA by the oracle's divsufsort shim), "
                  f"dist_anchor {walk_per_pair * 1e3:.2f} ms/pair; extrapolated to rows of {n_total - 1} queries",
        "esa_s_per_subject": esa_per_subject, "walk_ms_per_pair": walk_per_pair * 1e3, "wall_s": wall,
        # the checker's cells of this sample (popped before the line is printed): columns = `_ids`,
        # rows = its first entries
        "_cells": ref_cells, "_ids": list(subjects) + qids,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g, ln, lo, hi, seed, model = workload_of(args)
    from andi_b200 import synth

    threads = os.cpu_count() or 1
    d = divergences(g, lo, hi, seed)
    base = synth.base_genome(ln, seed)
    cache = {}

    def host_seq(i):
        if i not in cache:
            cache[i] = synth.ACGT[synth.mutate(base, float(d[i]), seed + 1 + i)].tobytes()
        return cache[i]

    S = min(threads, g - 1)
    results = []
    for step in range(args.warmup + args.steps):
        subjects = [(step * S + k) % g for k in range(S)]
        r = cpu_sample(host_seq, g, subjects, args.cpu_queries, model, threads)
        if step >= args.warmup:
            results.append(r)
        cache.clear()
    value = float(np.mean([r["value"] for r in results]))
    last = results[-1]
    last.pop("_cells", None), last.pop("_ids", None)
    last["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean([r["wall_s"] for r in results])) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {g} x {ln} bp star phylogeny d~U[{lo},{hi}], model {model}"},
        "cpu_baseline": last,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm

def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import ctypes as C
    import hashlib

    import torch
    import torch.distributed as dist

    from andi_b200 import driver, native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    store = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        # the atomic counter of the dynamic subject queue: add() of the rendezvous store the process
        # group already has; else a store of our own on the next port; else static row blocks
        try:
            store = dist.distributed_c10d._get_default_store()
            store.add("andi_probe", 1)
        except Exception:
            try:
                store = dist.TCPStore(os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29500")) + 1,
                                      world, is_master=(rank == 0))
            except Exception:
                store = None

    g, ln, lo, hi, seed, model = workload_of(args)
    chars, offsets, lens, d = make_pool_device(g, ln, lo, hi, seed, device, args.contigs, args.repeats)
    torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    L = native.load()
    ctx = native.Context(local, stream.cuda_stream)
    ctx.set_pool_device(chars.data_ptr(), offsets, lens)

    rows = args.rows or max(1, min(g, -(-(148 * 2048) // g)))  # one full walk batch
    rows = min(rows, g)
    out_dev = torch.empty((rows, g, 17), dtype=torch.int32, device=device)
    gathered = torch.empty((world * rows, g, 17), dtype=torch.int32, device=device) if world > 1 else None

    def first_row(step):
        return ((step * world + rank) * rows) % max(1, g - rows + 1)

    def step_device(step):
        s0 = first_row(step)
        ctx.dist_rows_device(out_dev.data_ptr(), s0, s0 + rows, 0.025, model)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- value: K steps with the pool resident in HBM, device-timed
    for w in range(args.warmup):
        step_device(w)
    barrier()
    ctx.reset_stats()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0.record(stream)
    for s in range(args.steps):
        step_device(args.warmup + s)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1))
    st = ctx.stats()
    pairs_per_step_rank = rows * (g - 1)
    value = world * pairs_per_step_rank * args.steps / (ms * 1e-3)

    # ---- the dominant kernel timed ALONE: one more step with a single subject in flight
    # (ANDI_B200_LANES=1), so that no other kernel of ours shares the SMs with the walk
    os.environ["ANDI_B200_LANES"] = "1"
    ctx.reset_stats()
    step_device(args.warmup + args.steps)
    barrier()
    st_alone = ctx.stats()
    del os.environ["ANDI_B200_LANES"]

    # ---- end to end through the host-buffer ABI: pinned host pool -> rows on the host. With
    # several ranks the pool is uploaded and packed once (rank 0) and its packed planes are
    # broadcast over NVLink; every rank computes its rows; rank 0 reads all of them back.
    e2e = None
    host_pool = None
    if not args.no_e2e:
        n = g
        ctx2 = native.Context(local, stream.cuda_stream)
        out_host = torch.empty((world * rows, g, 17), dtype=torch.int32, pin_memory=True) if rank == 0 else None
        if rank == 0:
            host_pool = torch.empty(chars.numel(), dtype=torch.uint8, pin_memory=True)
            host_pool.copy_(chars)
            torch.cuda.synchronize()
            base_ptr = host_pool.data_ptr()
            ptrs = (C.c_char_p * n)(*[C.c_char_p(base_ptr + o) for o in offsets])
            lens_c = (C.c_size_t * n)(*lens)

        def step_e2e(step):
            s0 = first_row(step)
            if rank == 0:
                ctx2._ck(L.andi_pool_set_host(ctx2.h, ptrs, lens_c, n))
                ctx2.n = n
            if world > 1:
                driver.broadcast_pool(ctx2, dist, device, rank)
                ctx2.dist_rows_device(out_dev.data_ptr(), s0, s0 + rows, 0.025, model)
                dist.all_gather_into_tensor(gathered, out_dev)
                if rank == 0:
                    out_host.copy_(gathered, non_blocking=True)
            else:
                ctx2._ck(L.andi_dist_rows(ctx2.h, s0, s0 + rows, 0.025, native.MODELS[model], 0, C.c_void_p(out_host.data_ptr())))

        for w in range(args.warmup):
            step_e2e(w)
        barrier()
        ctx2.reset_stats()
        e0.record(stream)
        for s in range(args.steps):
            step_e2e(args.warmup + s)
        e1.record(stream)
        barrier()
        ms2 = max_over_ranks(e0.elapsed_time(e1))
        st2 = ctx2.stats()
        d2h = int(st2["d2h_bytes"] // args.steps) if world == 1 else (world * rows * g * 68 if rank == 0 else 0)
        e2e = {
            "value": world * pairs_per_step_rank * args.steps / (ms2 * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": int(st2["h2d_bytes"] // args.steps), "d2h_bytes_per_step": d2h,
            "ms_per_step": ms2 / args.steps,
            "pool": "uploaded and packed on rank 0, packed planes broadcast with NCCL" if world > 1 else "uploaded and packed on the GPU",
        }
        ctx2.close()

    # ---- the whole matrix (strong scaling): all g rows, subjects from a shared queue in batches,
    # rows summed to rank 0 (disjoint, so the sum is the matrix)
    full_matrix = None
    if not args.no_full and (g * g * 68) < 8e9:
        full = torch.zeros((g, g, 17), dtype=torch.int32, device=device)
        batch = max(1, min(16, g // (world * 8)))
        local_next = [0]

        static_rows = driver.shard_subjects(g, world, rank) if (world > 1 and store is None) else None

        def take(b):
            if store is not None:
                return store.add("andi_next_subject", b) - b
            v = local_next[0] + (static_rows[0] if static_rows else 0)
            local_next[0] += b
            return v

        barrier()
        e0.record(stream)
        mine = driver.dynamic_rows(ctx, g, full.data_ptr(), take, batch, 0.025, model, limit=static_rows[1] if static_rows else None)
        if world > 1:
            dist.reduce(full, dst=0, op=dist.ReduceOp.SUM)
        e1.record(stream)
        barrier()
        ms3 = max_over_ranks(e0.elapsed_time(e1))
        counts = torch.tensor([mine], dtype=torch.int64, device=device)
        all_counts = [torch.zeros_like(counts) for _ in range(world)]
        if world > 1:
            dist.all_gather(all_counts, counts)
        else:
            all_counts = [counts]
        if rank == 0:
            digest = hashlib.blake2b(full.cpu().numpy().tobytes(), digest_size=16).hexdigest()
            full_matrix = {"seconds": ms3 * 1e-3, "pairs_per_s": g * (g - 1) / (ms3 * 1e-3), "rows": g, "queue_batch": batch,
                           "queue": "shared counter (store.add)" if store is not None else ("static row blocks" if world > 1 else "one rank"),
                           "rows_per_rank": [int(c.item()) for c in all_counts],
                           # the same digest at every N = the N-GPU matrix is the 1-GPU matrix
                           "blake2b_of_matrix": digest}
        del full

    # ---- roofline of the dominant kernel (the anchor walk), from the single-lane step
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = json.loads(peaks_file.read_text())["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bytes_per_pair = 2 * ln / 4  # query + subject diagonal, 2 bits per base, read once
    # one "launch" = the walk kernels of one subject (k_walk_v3<1>, k_walk_v3<2>, k_walk_reduce)
    n_walks = max(1, st_alone["subjects"])
    walk_ms = st_alone["walk_ms"] / n_walks
    pairs_per_launch = st_alone["pairs"] / n_walks
    achieved = pairs_per_launch * bytes_per_pair / (walk_ms * 1e-3) / 1e9 if walk_ms > 0 else 0.0
    traffic, traffic_src = None, None
    tf = ROOT / "profiles" / "walk_traffic.json"
    if tf.exists() and args.workload == "c4" and not args.genomes and not args.length and not args.contigs and not args.repeats:
        # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture (tools/capture_traffic.sh);
        # only a capture of THIS build of the kernels counts
        t_ = json.loads(tf.read_text())
        if t_.get("kernel_sha") == kernel_sources_sha():
            traffic = (t_["dram_bytes_read"] + t_["dram_bytes_write"]) * (pairs_per_launch / t_["pairs_per_launch"])
            traffic_src = t_.get("capture")
        else:
            traffic_src = "profiles/walk_traffic.json is from other kernel sources (%s): ignored" % t_.get("kernel_sha")
    roofline = {
        "bound": "hbm", "kernel": "k_walk_v3<1> + k_walk_v3<2> + k_walk_reduce (one subject)", "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "algorithmic_bytes_per_pair": bytes_per_pair, "pairs_per_launch": pairs_per_launch, "launch_ms": walk_ms,
        "timed": "alone, one subject in flight (ANDI_B200_LANES=1); the timed steps run two subjects in flight",
        "walk_share_of_step": (st_alone["walk_ms"] / st_alone["rows_ms"]) if st_alone["rows_ms"] > 0 else None,
    }
    esa_ms = st_alone["esa_ms"] / n_walks
    esa_bytes = 14 * (2 * ln + 1) + 16.8e6

    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        if host_pool is None:
            host_pool = chars.cpu()
        hp = host_pool.numpy()

        def host_seq(i):
            return hp[offsets[i] : offsets[i] + lens[i]].tobytes()

        threads = os.cpu_count() or 1
        S = max(1, min(threads, g - 1))
        cpu = cpu_sample(host_seq, g, list(range(S)), args.cpu_queries, model, threads)
        # ---- parity of the benched workload: the checker's cells of the sample against the cells
        # the timed context computes for the same (subject, query) pairs of the full pool
        ref_cells, ids = cpu.pop("_cells"), np.asarray(cpu.pop("_ids"))
        rows_chk = ref_cells.shape[0]
        chk = torch.empty((rows_chk, g, 17), dtype=torch.int32, device=device)
        ctx.dist_rows_device(chk.data_ptr(), 0, rows_chk, 0.025, model)
        torch.cuda.synchronize()
        got = chk.cpu().numpy().view(np.uint32)[:, ids, :]
        bad = np.argwhere((got != ref_cells).any(axis=2))
        parity = {"cells": int(got.shape[0] * got.shape[1]), "mismatches": int(len(bad)), "against": cpu["kind"],
                  "subjects": rows_chk, "queries_per_subject": int(len(ids))}
        if len(bad):
            parity["first_bad"] = [[int(ids[r]), int(ids[c])] for r, c in bad[:5]]

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {
                "workload": f"{args.workload}: {g} x {ln} bp star phylogeny d~U[{lo},{hi}], model {model}; "
                            f"step = {rows} subject rows x {g - 1} queries per GPU (index build + walk)",
                "rows_per_step_per_gpu": rows, "pairs_per_step": world * pairs_per_step_rank,
                "l2": "inputs larger than L2 (packed pool %.0f MB, different subjects every step)" % (g * ln / 4 / 1e6),
                "parallelism": f"subjects sharded over {world} GPU(s), pool replicated, rows all-gathered (NCCL)" if world > 1 else "single GPU",
            },
            "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(st["esa_launches"] + st["walk_launches"]),
            "cub_calls": int(st["cub_calls"]),
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "full_matrix": full_matrix,
            "esa_build": {"mbp_per_s": ln / 1e6 / (esa_ms * 1e-3) if esa_ms > 0 else None, "ms_per_subject": esa_ms,
                          "algorithmic_gbs": (esa_bytes / (esa_ms * 1e-3) / 1e9) if esa_ms > 0 else None,
                          "sa_rounds_per_subject": st["sa_rounds"] / max(1, st["subjects"]),
                          "timed": "alone (single lane); in the timed steps it overlaps the previous subject's walk"},
            "kernel_ms_sums_of_timed_steps": {"walk": st["walk_ms"], "esa": st["esa_ms"], "rows_wall": st["rows_ms"]},
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    if parity and parity["mismatches"]:
        raise SystemExit(f"bench.py: the benched workload differs from the {parity['against']} in {parity['mismatches']} cells")


def kernel_sources_sha():
    """Digest of the kernel sources: ties a profile capture to the code it was taken from."""
    import hashlib

    h = hashlib.sha1()
    for f in sorted((ROOT / "andi_b200" / "csrc").glob("*.cu*")) + sorted((ROOT / "andi_b200" / "csrc").glob("*.h")):
        h.update(f.read_bytes())
    return h.hexdigest()[:12]


if __name__ == "__main__":
    main()
