#!/usr/bin/env python
"""bench.py — read pairs mapped / s on B200 for VirGenA's read-mapping hot path.

Contract: `python bench.py --gpus N --steps K --warmup W` (N>1 under torchrun, one rank per GPU) prints
ONE JSON line from rank 0. A "step" is one pass of the hot path — Mapper.mapReads (seeding, pair
selection, Smith-Waterman fill + traceback) followed by ConsensusBuilderSimple.getConsensusSimple
(pileup, all-reduce of the counts over ranks, argmax) — over one batch of synthetic read pairs.

Workload: BASELINE.json configs[2] ("synthetic 10M 2x250 HIV-1 read pairs at 15-25 % divergence", the
largest single-GPU configuration that exercises the SW/GCUPS half of the metric), processed as batches of
--pairs (default 1,000,000) pairs per step per GPU; every rank maps its own shard (weak scaling), the
only collective is the NCCL all-reduce of the int32 pileup counts.

  value  : pairs/s with the reads already resident in HBM when the timed region starts (CUDA events on
           the library's stream, max over ranks).
  e2e    : the same through the reference-facing C-ABI calls with HOST (pinned) buffers: vg_upload_reads
           (H2D) + vg_map_pairs + vg_pileup + vg_consensus + vg_get_results (D2H) inside the timed region.
  roofline, cpu_baseline: see DESIGN.md §Measurement.

`--impl reference` times the oracle (C++ restatement of the Java mapper; the JVM reference cannot run
here) on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

METRIC = "read_pairs_mapped_per_s"
UNIT = "pairs/s"
WORKLOAD = "cfg3: synthetic 2x250 HIV-1 (9719 bp) read pairs, 15-25% divergence, 1% indels; K=5, 2/-3/5/2"


def load_cutoffs(cfg):
    with open(os.path.join(ROOT, "tests", "golden", "cutoffs.json")) as f:
        d = json.load(f)["cfg%d" % cfg]
    return np.array(d["cutoffs"], np.int32), d


def gen_pairs_torch(ref_np, n_pairs, L, seed, device, sub_lo=0.15, sub_hi=0.25, indel_rate=0.01):
    """Config-3 reads generated on the device (torch is plumbing here: the numpy generator of
    virgena_b200.synth is too slow for 10^6 pairs per batch). Returns uint8 [2, n_pairs, L]."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ref = torch.from_numpy(ref_np).to(device)
    G = ref.numel()
    acgt = torch.tensor(list(b"ATGC"), dtype=torch.uint8, device=device)
    code = torch.zeros(256, dtype=torch.long, device=device)
    code[acgt.long()] = torch.arange(4, device=device)
    comp = torch.zeros(256, dtype=torch.uint8, device=device)
    for a, b in (b"AT", b"TA", b"GC", b"CG", b"NN"):
        comp[a] = b
    frag = (torch.randn(n_pairs, generator=g, device=device) * 80 + 400).long().clamp(L, 1000)
    start = (torch.rand(n_pairs, generator=g, device=device) * (G - frag + 1).float()).long()
    sub = torch.rand(n_pairs, generator=g, device=device) * (sub_hi - sub_lo) + sub_lo

    def sample(st):
        u = torch.rand(n_pairs, L, generator=g, device=device)
        adv = torch.ones(n_pairs, L, dtype=torch.long, device=device)
        dele = u < indel_rate / 2
        ins = (u >= indel_rate / 2) & (u < indel_rate)
        adv = adv + dele.long() * torch.randint(1, 4, (n_pairs, L), generator=g, device=device)
        adv = torch.where(ins, torch.zeros_like(adv), adv)
        idx = st[:, None] + torch.cumsum(adv, 1) - adv[:, :1]
        idx = idx.clamp(0, G - 1)
        r = ref[idx]
        rnd = acgt[torch.randint(0, 4, (n_pairs, L), generator=g, device=device)]
        r = torch.where(ins, rnd, r)
        m = torch.rand(n_pairs, L, generator=g, device=device) < sub[:, None]
        shifted = acgt[(code[r.long()] + torch.randint(1, 4, (n_pairs, L), generator=g, device=device)) % 4]
        return torch.where(m, shifted, r)

    left = sample(start)
    right = comp[sample((start + frag - L).clamp(min=0)).long()].flip(1)
    flip = torch.rand(n_pairs, generator=g, device=device) < 0.5
    m1 = torch.where(flip[:, None], right, left)
    m2 = torch.where(flip[:, None], left, right)
    return torch.stack([m1, m2]).contiguous()


class ClockSampler:
    def __init__(self, dev_index):
        self.dev = dev_index
        self.rows = []
        self.stop = False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.t.start()

    def finish(self):
        self.stop = True
        self.t.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_baseline(ref, cut, m1, m2, n_sample, threads, index_threads):
    """The oracle ("port": C++ restatement of the Java mapper) on the host cores, bounded sample."""
    import oracle
    from virgena_b200 import synth
    o = oracle.Oracle(K=5, cutoffs=cut)
    o.set_reference(ref, index_thread_num=index_threads)
    b = synth.pack_pairs(m1[:n_sample], m2[:n_sample])
    t0 = time.time()
    rec, pool, st = o.map_pairs(*b, batch_size=1000, n_threads=threads)
    counts = oracle.pileup(rec, pool, len(ref))
    oracle.consensus(counts, ref, True, 0)
    dt = time.time() - t0
    return n_sample / dt, dt


def msa_arm(vg, synth, device, n_pairs):
    """Secondary line: BASELINE.json configs[1], MapperMSA.mapAllReads (seeding + pair selection against the reference
    MSA, no alignment) on synthetic cfg2 reads, device-resident, CUDA events inside the library. Rank 0 only."""
    cut, _ = load_cutoffs(2)
    rows = synth.config_msa()
    m1, m2 = synth.config_pairs(2, min(n_pairs, 5000))
    k = (n_pairs + len(m1) - 1) // len(m1)
    m1 = np.tile(m1, (k, 1))[:n_pairs]
    m2 = np.tile(m2, (k, 1))[:n_pairs]
    aln = vg.ReferenceAlignment(rows, K=7)
    mapper = vg.MapperMSA(K=7, indel_tolerance=1.5, cutoffs=cut, device=device)
    reads = vg.PairedReads(*synth.pack_pairs(m1, m2))
    best = None
    for _ in range(4):
        md = mapper.mapAllReads(reads, aln)
        t = mapper.ctx.timings()["total_ms"]
        best = t if best is None else min(best, t)
    sd = mapper.ctx.seed_timings()
    return {"metric": "read_pairs_mapped_per_s (MapperMSA.mapAllReads)", "value": n_pairs / (best * 1e-3), "unit": UNIT,
            "ms": best, "pairs": n_pairs, "workload": "cfg2: synthetic 2x250 reads against a %d-sequence MSA (%d columns), K=7, 1.5"
            % (rows.shape[0], rows.shape[1]), "seed_gen_msa_kernel_ms": sd["gen_ms"], "seed_win_kernel_ms": sd["win_ms"],
            "fallback_items": int(sd["fallback_items"]), "mapping_stats": [int(x) for x in md.stats]}


def run_reference(args):
    """--impl reference: the CPU path on all host cores. Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from virgena_b200 import synth
    cut, meta = load_cutoffs(3)
    ref, ends = synth.config_reference(3)
    threads = os.cpu_count() or 1
    n_sample = args.ref_pairs or 1000 * threads  # BatchSize=1000 pairs per task: one task per thread
    m1, m2 = synth.config_pairs(3, n_sample)
    for _ in range(args.warmup):
        cpu_baseline(ref, cut, m1, m2, min(n_sample, 200), threads, meta["index_thread_num"])
    t0 = time.time()
    for _ in range(args.steps):
        cpu_baseline(ref, cut, m1, m2, n_sample, threads, meta["index_thread_num"])
    dt = time.time() - t0
    v = n_sample * args.steps / dt
    sample = "%d pairs of the workload per step on %d host threads (BatchSize=1000 tasks, like Mapper.java:436-445)" % (n_sample, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_step": n_sample, "read_len": 250},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C++ restatement of the reference (oracle/), not the JVM: no JDK in this environment"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--pairs", type=int, default=1000000, help="read pairs per step per GPU")
    ap.add_argument("--ref-pairs", type=int, default=0, help="pairs per step of the CPU reference arm (0: 1000 x host threads)")
    ap.add_argument("--cpu-pairs", type=int, default=0, help="pairs of the cpu_baseline sample (0: 3000 x host threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-msa", action="store_true", help="skip the secondary MapperMSA.mapAllReads measurement (configs[1])")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import virgena_b200 as vg
    from virgena_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    L = 250
    cut, meta = load_cutoffs(3)
    ref, ends = synth.config_reference(3)
    G = len(ref)
    n = args.pairs
    ctx = vg.Context(K=5, low_coef=0.0, high_coef=1.25, cutoffs=cut, insert_len=1000, device=local)
    genome = vg.Reference(ref, K=5, thread_num=meta["index_thread_num"])
    ctx.set_reference(genome)
    ext_stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    # one synthetic batch per step would need 0.5 GB each; two distinct batches are alternated (each is
    # far larger than the 126 MB L2, so nothing survives in cache between steps)
    n_batches = 2
    host_batches = []
    for b in range(n_batches):
        t = gen_pairs_torch(ref, n, L, synth.SEED0 + 3 + 1000 * (rank * n_batches + b + 1), dev)
        hb = torch.empty(t.shape, dtype=torch.uint8, pin_memory=True)
        hb.copy_(t)
        host_batches.append(hb)
        del t
    torch.cuda.synchronize()
    def pinned(a):  # host arrays of the e2e arm live in pinned memory (no staging copies inside cudaMemcpyAsync)
        t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
        t.copy_(torch.from_numpy(a))
        return t.numpy()
    off1 = pinned(np.arange(n, dtype=np.int64) * L)
    off2 = pinned(n * L + off1)
    len1 = pinned(np.full(n, L, np.int32))
    len2 = pinned(np.full(n, L, np.int32))
    n_bases = 2 * n * L

    counts_tensor = [None]

    def allreduce_counts(c=None):
        if world == 1:
            return
        p, cnt = (c or ctx).counts_device()

        class _W:
            __cuda_array_interface__ = {"shape": (cnt,), "typestr": "<i4", "data": (p, False), "version": 3}
        tns = torch.as_tensor(_W(), device=dev)
        dist.all_reduce(tns)
        torch.cuda.current_stream().synchronize()
        counts_tensor[0] = tns

    def upload(b):
        ctx.upload_reads_raw(host_batches[b].data_ptr(), n_bases, off1, len1, off2, len2)

    def step_resident():
        st = ctx.map_pairs(fetch=False)
        ctx.pileup(False)
        allreduce_counts()
        ctx.consensus(ref, True, 0)
        return st

    # e2e arm: two host threads, each with its own context (= its own stream and device buffers), take the steps in
    # turn, so that one thread's PCIe copies run while the other's kernels do — the way a multi-threaded host (the
    # reference maps with a thread pool, Mapper.java:436-445) drives one context per thread. Every step still does its
    # own H2D of the reads and D2H of the results inside the timed region; the all-reduces are issued in step order.
    n_e2e = 2
    e2e_ctx = [ctx]
    for _ in range(n_e2e - 1):
        c2 = vg.Context(K=5, low_coef=0.0, high_coef=1.25, cutoffs=cut, insert_len=1000, device=local)
        c2.set_reference(genome)
        e2e_ctx.append(c2)
    rec_hosts = [torch.zeros(n * vg.PAIR_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True).numpy().view(vg.PAIR_DTYPE) for _ in range(n_e2e)]
    pool_hosts = [torch.empty(n * 2 * (L + L * 3 // 2 + 8), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(n_e2e)]
    d2h = [0]
    turn = [0]
    turn_cv = threading.Condition()

    def step_e2e(tid, b, seq):
        c = e2e_ctx[tid]
        c.upload_reads_raw(host_batches[b].data_ptr(), n_bases, off1, len1, off2, len2)
        st = c.map_pairs(fetch=False)
        nres, nbytes = c.get_results_into(rec_hosts[tid], pool_hosts[tid])
        c.pileup(False)
        with turn_cv:  # collectives in the same order on every rank
            while turn[0] < seq:
                turn_cv.wait()
            if turn[0] != seq:
                raise RuntimeError("another e2e worker failed")
        if world > 1:
            torch.cuda.set_device(local)
        allreduce_counts(c)
        with turn_cv:
            turn[0] = seq + 1
            turn_cv.notify_all()
        cons = c.consensus(ref, True, 0)
        d2h[0] = nres * vg.PAIR_DTYPE.itemsize + nbytes + len(cons)
        return st

    def run_e2e(n_steps):
        turn[0] = 0
        errs = []

        def worker(tid):
            try:
                for sq in range(tid, n_steps, n_e2e):
                    step_e2e(tid, sq % n_batches, sq)
            except BaseException as e:  # noqa: BLE001
                errs.append(e)
                with turn_cv:
                    turn[0] = 1 << 30
                    turn_cv.notify_all()
        ths = [threading.Thread(target=worker, args=(t,)) for t in range(n_e2e)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errs:
            raise errs[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ---------------------------------------------------------------------------
    launches0 = ctx.kernel_launches()
    upload(0)
    for w in range(args.warmup):
        stats = step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    stage_ms = np.zeros(6)
    seed_acc = np.zeros(8)
    barrier()
    launches1 = ctx.kernel_launches()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record(ext_stream)
    for s in range(args.steps):
        stats = step_resident()
        t = ctx.timings()
        stage_ms += np.array([t["seed_ms"], t["sw_fill_ms"], t["traceback_ms"], t["pileup_ms"], t["total_ms"], t["sw_cells"]])
        sd = ctx.seed_timings()
        seed_acc += np.array([sd["gen_ms"], sd["win_ms"], sd["fallback_ms"], sd["longhits_in"], sd["longhits_kept"], sd["fallback_items"],
                              sd["windows_out"], sd["items"]])
    ev1.record(ext_stream)
    barrier()
    wall_ms = (time.time() - t0) * 1e3
    dev_ms = ev0.elapsed_time(ev1)
    launches2 = ctx.kernel_launches()
    clocks = sampler.finish() if rank == 0 else None
    # the step also contains host-side waits (counter read-backs); report the larger of the two clocks
    ms = max(dev_ms, wall_ms)
    tt = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    value = n * world * args.steps / (ms * 1e-3)

    # ---- end-to-end arm ----------------------------------------------------------------------------------
    run_e2e(2 * n_e2e)
    barrier()
    t0 = time.time()
    run_e2e(args.steps)
    barrier()
    e2e_ms = (time.time() - t0) * 1e3
    tt = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_ms = float(tt.item())
    e2e_value = n * world * args.steps / (e2e_ms * 1e-3)
    h2d = n_bases + off1.nbytes * 2 + len1.nbytes * 2

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        seed_ms, fill_ms, tb_ms, pu_ms, tot_ms, cells = stage_ms / args.steps
        dpx = ctx.dpx_peak()
        # The seeding stage is two kernels: seed_gen_kernel (reads -> LongHit stream, 8 B per LongHit)
        # and seed_win_kernel (stream -> windows). Algorithmic HBM bytes per launch:
        #   gen: read bases in (every read is seeded in both orientations: 2 x (L1 + L2) per pair) + stream out
        #   win: stream in + window records out (12 B per window + 4 B count per orientation)
        gen_ms, win_ms, fb_ms, lh_in, lh_kept, fb_items, win_out, items = seed_acc / args.steps
        gen_bytes = 4 * L * n + 8 * lh_in
        win_bytes = 8 * lh_in + 12 * win_out + 4 * items
        kern = {"seed_gen_kernel": (gen_ms, gen_bytes), "seed_win_kernel": (win_ms, win_bytes)}
        traffic_all = {}
        tp = os.path.join(ROOT, "profiles", "traffic_r1.json")
        if os.path.exists(tp):
            traffic_all = json.load(open(tp)).get("dram_bytes_per_pair", {})
        rl = {}
        for k, (kms, kb) in kern.items():
            gbs = kb / (kms * 1e-3) / 1e9 if kms else None
            rl[k] = {"bound": "hbm", "kernel": k, "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": gbs / hbm_peak if gbs else None,
                     "traffic": traffic_all[k] * n if k in traffic_all else None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": kb, "ms_per_step": kms}
        dominant = max(kern, key=lambda k: kern[k][0])
        # pileup: algorithmic bytes = aligned columns read once (1 B each) + 16 B per mapped read + G*5*4 B out
        n_res, seq1_bytes = ctx.result_sizes()
        pile_bytes = seq1_bytes + 16 * int(stats[3]) + G * 5 * 4
        pile = {"bound": "hbm", "achieved": pile_bytes / (pu_ms * 1e-3) / 1e9 if pu_ms else None, "peak": hbm_peak,
                "unit": "GB/s", "frac": pile_bytes / (pu_ms * 1e-3) / 1e9 / hbm_peak if pu_ms else None,
                "note": "all pileup kernels (list, bin, count, reduce); 5 ms per 10^6 pairs"}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step_per_gpu": n, "read_len": L, "reference_len": G,
                       "cache": "inputs larger than L2 (0.5 GB of reads per step, two alternating batches)",
                       "step": "mapReads + getConsensusSimple (seed, pair-select, SW fill, traceback, pileup, allreduce, argmax)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h[0]),
                    "ms_per_step": e2e_ms / args.steps,
                    "how": "%d host threads with one context each alternate over the steps (copies of one overlap kernels of the "
                           "other); each step = vg_upload_reads + vg_map_pairs + vg_get_results + vg_pileup + vg_consensus" % n_e2e},
            "gpu_launches": int(launches2 - launches1),
            "clocks": clocks,
            "stage_ms_per_step": {"seed_pairselect": seed_ms, "seed_gen_kernel": gen_ms, "seed_win_kernel": win_ms,
                                  "seed_fallback_kernel": fb_ms, "sw_fill": fill_ms, "sw_traceback": tb_ms, "pileup": pu_ms,
                                  "map_total": tot_ms},
            "roofline": dict(rl[dominant], note="dominant kernel of the step; the seeding kernels are issue / shared-memory bound by "
                             "construction (SURVEY 8d): LongHits generated per second = %.3e, swept per second = %.3e; "
                             "%d of %d read orientations took the general (fallback) kernel"
                             % (lh_in / (gen_ms * 1e-3) if gen_ms else 0, lh_in / (win_ms * 1e-3) if win_ms else 0, int(fb_items), int(items))),
            "seed_gen_roofline": rl["seed_gen_kernel"],
            "seed_win_roofline": rl["seed_win_kernel"],
            "sw_roofline": {"bound": "dpx", "gcups_fill": cells / (fill_ms * 1e-3) / 1e9 if fill_ms else None,
                            "gcups_fill_traceback": cells / ((fill_ms + tb_ms) * 1e-3) / 1e9 if fill_ms else None,
                            "dpx_s16x2_ops_per_s_measured": dpx,
                            "gcups_peak_3ops_per_cellpair": dpx * 2 / 3 / 1e9,
                            "frac_of_3op_peak_fill": (cells / (fill_ms * 1e-3)) / (dpx * 2 / 3) if fill_ms else None,
                            "dpx_ops_per_cellpair_issued": 7.5,
                            "frac_of_dpx_issue_peak_fill": (cells / 2 * 7.5 / (fill_ms * 1e-3)) / dpx if fill_ms else None,
                            "cells_per_step": cells},
            "pileup_roofline": pile,
            "mapping_stats": [int(x) for x in stats],
        }
        if not args.no_msa:
            out["msa_mode"] = msa_arm(vg, synth, local, min(n, 200000))
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            ncpu = min(n, args.cpu_pairs or 3000 * threads)
            m1 = host_batches[0][0][:ncpu].numpy()
            m2 = host_batches[0][1][:ncpu].numpy()
            v, dt = cpu_baseline(ref, cut, m1, m2, ncpu, threads, meta["index_thread_num"])
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                   "sample": "first %d pairs of batch 0 (%.1f s wall on %d threads, BatchSize=1000 tasks)" % (ncpu, dt, threads),
                                   "note": "C++ restatement of the reference (oracle/), not the JVM"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
