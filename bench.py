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


def cpu_baseline(ref, cut, m1, m2, n_sample, threads, index_threads, want_results=False, contig_ends=None):
    """The oracle ("port": C++ restatement of the Java mapper) on the host cores, bounded sample."""
    import oracle
    from virgena_b200 import synth
    o = oracle.Oracle(K=5, cutoffs=cut)
    frag = contig_ends is not None and len(contig_ends) > 1
    o.set_reference(ref, contig_ends=contig_ends, fragment_ends=contig_ends, is_fragmented=frag, index_thread_num=index_threads)
    b = synth.pack_pairs(m1[:n_sample], m2[:n_sample])
    t0 = time.time()
    rec, pool, st = o.map_pairs(*b, batch_size=1000, n_threads=threads)
    counts = oracle.pileup(rec, pool, len(ref))
    cons = oracle.consensus(counts, ref, True, 0)
    dt = time.time() - t0
    if want_results:
        return n_sample / dt, dt, (rec, pool, st, counts, cons)
    return n_sample / dt, dt


PARITY_FIELDS = ("present", "start", "end", "count", "reverse", "fragment_id", "score", "start1", "end1", "start2", "end2",
                 "identity", "length", "seq1_offset")


def parity_check(ctx, genome_ref, reads_packed, oracle_out):
    """Outside every timed region: the device maps the very pairs the oracle mapped for cpu_baseline and the two results
    are compared field by field (records, sequence1 bytes, the nine statistics, pileup counts, consensus)."""
    import virgena_b200 as vg
    orec, opool, ostats, ocounts, ocons = oracle_out
    ctx.upload_reads(vg.PairedReads(*reads_packed))
    md = ctx.map_pairs()
    ctx.pileup(False)
    cons = ctx.consensus(genome_ref, True, 0)
    bad = []
    if not (md.records["cls"] == orec["cls"]).all():
        bad.append("class")
    for f in PARITY_FIELDS:
        a, b = md.records["m"][f], orec["m"][f]
        if f == "seq1_offset":
            pres = orec["m"]["present"] == 1
            a, b = a[pres], b[pres]
        if not (a == b).all():
            bad.append(f)
    if md.seq1_pool.tobytes() != opool.tobytes():
        bad.append("sequence1")
    if not (md.stats == ostats).all():
        bad.append("stats")
    if not (ctx.counts() == ocounts).all():
        bad.append("pileup counts")
    if cons.tobytes() != ocons.tobytes():
        bad.append("consensus")
    return bad


def msa_arm(vg, synth, device, n_pairs):
    """Secondary line: BASELINE.json configs[1], MapperMSA.mapAllReads (seeding + pair selection against the reference
    MSA, no alignment) on synthetic cfg2 reads, device-resident, CUDA events inside the library. Rank 0 only."""
    cut, _ = load_cutoffs(2)
    rows = synth.config_msa()
    m1, m2 = synth.config_pairs(2, n_pairs)  # all pairs distinct (0.5 GB of reads per 10^6 pairs: nothing fits in L2)
    aln = vg.ReferenceAlignment(rows, K=7)
    mapper = vg.MapperMSA(K=7, indel_tolerance=1.5, cutoffs=cut, device=device)
    reads = vg.PairedReads(*synth.pack_pairs(m1, m2))
    ts = []
    for _ in range(4):  # one warm-up + three timed calls: the median is reported
        md = mapper.mapAllReads(reads, aln)
        ts.append(mapper.ctx.timings()["total_ms"])
    best = float(np.median(ts[1:]))
    sd = mapper.ctx.seed_timings()
    return {"metric": "read_pairs_mapped_per_s (MapperMSA.mapAllReads)", "value": n_pairs / (best * 1e-3), "unit": UNIT,
            "ms": best, "pairs": n_pairs, "workload": "cfg2: synthetic 2x250 reads against a %d-sequence MSA (%d columns), K=7, 1.5"
            % (rows.shape[0], rows.shape[1]), "seed_gen_msa_kernel_ms": sd["gen_ms"], "seed_win_kernel_ms": sd["win_ms"],
            "fallback_items": int(sd["fallback_items"]), "mapping_stats": [int(x) for x in md.stats]}


def secondary_arm(vg, synth, device, cfg, n_pairs, threads_cpu, check_pairs):
    """Secondary line for BASELINE.json configs[3] (cfg 4: 8-segment influenza-like genome, 2x150, fragmented) and configs[4]
    (cfg 5: six-genotype HCV mixture mapped to genotype 1, 2x150): mapReads + getConsensusSimple, device-resident, timed by
    the library's CUDA events, best of 3; the first `check_pairs` pairs are also mapped by the oracle and compared. Rank 0."""
    cut, meta = load_cutoffs(cfg)
    ref, ends = synth.config_reference(cfg)
    base = min(n_pairs, 100000)  # the numpy generator makes 10^5 distinct pairs; more are distinct batches of them
    parts1, parts2 = [], []
    for k in range((n_pairs + base - 1) // base):
        a, b = synth.config_pairs(cfg, base, seed_offset=k)
        parts1.append(a)
        parts2.append(b)
    m1 = np.concatenate(parts1)[:n_pairs]
    m2 = np.concatenate(parts2)[:n_pairs]
    frag = len(ends) > 1
    ctx = vg.Context(K=5, low_coef=0.0, high_coef=1.25, cutoffs=cut, insert_len=1000, device=device)
    genome = vg.Reference(ref, K=5, contig_ends=ends, fragment_ends=ends, is_fragmented=frag, thread_num=meta["index_thread_num"])
    ctx.set_reference(genome)
    ctx.upload_reads(vg.PairedReads(*synth.pack_pairs(m1, m2)))
    best = None
    for _ in range(4):
        st = ctx.map_pairs(fetch=False)
        ctx.pileup(False)
        ctx.consensus(ref, True, 0)
        t = ctx.timings()
        tot = t["total_ms"] + t["pileup_ms"]
        if best is None or tot < best[0]:
            best = (tot, t, ctx.seed_timings())
    tot, t, sd = best
    out = {"metric": METRIC, "value": n_pairs / (tot * 1e-3), "unit": UNIT, "ms": tot, "pairs": n_pairs,
           "workload": "cfg%d: %s" % (cfg, "synthetic 8-segment influenza-A-like genome (13588 bp), 2x150, 3% divergence, fragmented"
                                      if cfg == 4 else "synthetic HCV-like six-genotype mixture (~30% mutual divergence) mapped to genotype 1 (9600 bp), 2x150"),
           "stage_ms": {"seed_pairselect": t["seed_ms"], "seed_gen_kernel": sd["gen_ms"], "seed_win_kernel": sd["win_ms"],
                        "seed_fallback_kernel": sd["fallback_ms"], "sw_fill": t["sw_fill_ms"], "sw_traceback": t["traceback_ms"],
                        "pileup": t["pileup_ms"]},
           "gcups_fill": t["sw_cells"] / (t["sw_fill_ms"] * 1e-3) / 1e9 if t["sw_fill_ms"] else None,
           "fallback_items": int(sd["fallback_items"]), "mapping_stats": [int(x) for x in st]}
    if check_pairs:
        nck = min(check_pairs, n_pairs)
        v, dt, res = cpu_baseline(ref, cut, m1, m2, nck, threads_cpu, meta["index_thread_num"], want_results=True, contig_ends=ends)
        bad = parity_check(ctx, ref, synth.pack_pairs(m1[:nck], m2[:nck]), res)
        out["parity_checked_pairs"] = nck
        out["parity_ok"] = not bad
        if bad:
            out["parity_mismatch"] = bad
        out["cpu_pairs_per_s"] = v
    ctx.close()
    return out


def random_model_arm(vg, synth, device, check):
    """Secondary line: the cutoffs of config 1 from the random model (RandomModel.genModel + computeCuttofs,
    10^4 Markov reads of 250 bases generated and scored on the device); with `check` the oracle's seeded restatement derives
    the same cutoffs on one host core and the two are compared. Rank 0, outside every timed region."""
    ref, ends = synth.config_reference(1)
    genome = vg.Reference(ref, K=5, thread_num=8)
    rm = vg.RandomModel(genome, K=5, order=4, coef=1.25, seed=20260102, device=device)
    rm.cutoffs(readNum=2000)  # warm-up (buffers, kernels)
    t0 = time.perf_counter()
    cut = rm.cutoffs(readNum=10000, minReadLen=250, maxReadLen=250, step=10, pValue=0.01)
    dt = time.perf_counter() - t0
    out = {"what": "RandomModel.genModel(10^4 reads x 250 bases) + computeCuttofs through vg_random_model_counts, wall time of the "
                   "call incl. the host-side sort", "ms": dt * 1e3, "reads_per_s": 10000 / dt, "cutoff_250": int(cut[250])}
    if check:
        import oracle
        t0 = time.perf_counter()
        ocut = oracle.cutoffs(ref, K=5, coef=1.25, order=4, read_num=10000, min_len=250, max_len=250, step=10, p_value=0.01,
                              index_thread_num=8, seed=20260102)
        out["cpu_port_ms_1_core"] = (time.perf_counter() - t0) * 1e3
        out["parity_ok"] = bool((ocut == cut).all())
    rm.ctx.close()
    return out


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
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary cfg4 / cfg5 lines (configs[3], configs[4])")
    ap.add_argument("--secondary-pairs", type=int, default=200000, help="pairs of each secondary cfg4 / cfg5 line")
    ap.add_argument("--check-pairs", type=int, default=20000, help="pairs of the sharded-vs-single check at N > 1")
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

    # The data path has no host-language collective: the all-reduce of the pileup counts is vg_pileup_allreduce (NCCL inside
    # the library, on the context's own stream). torch.distributed only carries the bootstrap (the 128-byte NCCL id per
    # communicator), the barriers and the max-over-ranks of the timings.
    def make_comm(c):
        if world == 1:
            return
        ids = [vg.Context.comm_unique_id().tobytes() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        c.comm_init_rank(np.frombuffer(ids[0], np.uint8), rank, world)

    make_comm(ctx)
    allreduce_ms = []

    def allreduce_counts(c=None):
        (c or ctx).pileup_allreduce()
        if c is None and world > 1:
            allreduce_ms.append(ctx.allreduce_ms())

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
        make_comm(c2)
        e2e_ctx.append(c2)
    rec_hosts = [torch.zeros(n * vg.PAIR_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True).numpy().view(vg.PAIR_DTYPE) for _ in range(n_e2e)]
    pool_hosts = [torch.empty(n * 2 * (L + L * 3 // 2 + 8), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(n_e2e)]
    ops_hosts = [torch.empty(n * 2 * 24, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32) for _ in range(n_e2e)]
    # the packed transfer form of the same batches (2 bits per ACGT base + exceptions), made once by the library's host-side
    # packer, in pinned memory: the form a read loader hands to vg_upload_reads_packed
    import ctypes as C
    packed_batches = []
    for b in range(n_batches):
        hb = host_batches[b].numpy().reshape(-1)
        pk = torch.zeros((n_bases + 3) // 4 + 64, dtype=torch.uint8, pin_memory=True)
        cap = max(1 << 16, n_bases // 50)
        epos = torch.zeros(cap, dtype=torch.int64, pin_memory=True)
        ebyt = torch.zeros(cap, dtype=torch.uint8, pin_memory=True)
        ne = C.c_int64()
        rc = vg._lib.lib().vg_pack_reads(C.c_void_p(hb.ctypes.data), n_bases, C.c_void_p(pk.data_ptr()), C.c_void_p(epos.data_ptr()),
                                         C.c_void_p(ebyt.data_ptr()), cap, C.byref(ne))
        if rc != 0:
            raise RuntimeError("vg_pack_reads: %d (%d exceptions)" % (rc, ne.value))
        packed_batches.append((pk, epos.numpy()[:ne.value], ebyt.numpy()[:ne.value]))
    d2h = [0]
    turn = [0]
    turn_cv = threading.Condition()
    e2e_mode = ["packed"]

    def step_e2e(tid, b, seq):
        c = e2e_ctx[tid]
        if e2e_mode[0] == "packed":
            pk, epos, ebyt = packed_batches[b]
            c.upload_reads_packed_raw(pk.data_ptr(), n_bases, epos, ebyt, off1, len1, off2, len2)
            st = c.map_pairs(fetch=False)
            nres = n
            nbytes = 4 * c.get_results_compact_into(rec_hosts[tid], ops_hosts[tid])
        else:
            c.upload_reads_raw(host_batches[b].data_ptr(), n_bases, off1, len1, off2, len2)
            st = c.map_pairs(fetch=False)
            nres, nbytes = c.get_results_into(rec_hosts[tid], pool_hosts[tid])
        c.pileup(False)
        with turn_cv:  # collectives in the same order on every rank
            while turn[0] < seq:
                turn_cv.wait()
            if turn[0] != seq:
                raise RuntimeError("another e2e worker failed")
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

    # ---- end-to-end arms -----------------------------------------------------------------------------------
    # headline: packed upload (vg_upload_reads_packed) + compact results (vg_get_results_compact); beside it the byte API
    # of round 1 (vg_upload_reads + vg_get_results with the full sequence1 pool)
    def time_e2e(mode):
        e2e_mode[0] = mode
        run_e2e(2 * n_e2e)
        barrier()
        t0 = time.time()
        run_e2e(args.steps)
        barrier()
        ms_ = (time.time() - t0) * 1e3
        tt_ = torch.tensor([ms_], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt_, op=dist.ReduceOp.MAX)
        return float(tt_.item()), int(d2h[0])

    bytes_ms, bytes_d2h = time_e2e("bytes")
    bytes_value = n * world * args.steps / (bytes_ms * 1e-3)
    bytes_h2d = n_bases + off1.nbytes * 2 + len1.nbytes * 2
    e2e_ms, e2e_d2h = time_e2e("packed")
    e2e_value = n * world * args.steps / (e2e_ms * 1e-3)
    h2d = (n_bases + 3) // 4 + 9 * int(np.mean([len(p[1]) for p in packed_batches])) + off1.nbytes * 2 + len1.nbytes * 2
    d2h[0] = e2e_d2h

    # ---- N > 1: the sharded path against the single-GPU result (not timed) --------------------------------------------
    multi_check = None
    if world > 1:
        nck = max(world, args.check_pairs)
        from virgena_b200.dist import shard_range
        ck1, ck2 = synth.config_pairs(5, nck)   # config 5 (configs[4]: the sharded one): the same fixed set on every rank
        cut5, meta5 = load_cutoffs(5)
        ref5, ends5 = synth.config_reference(5)
        cctx = vg.Context(K=5, low_coef=0.0, high_coef=1.25, cutoffs=cut5, insert_len=1000, device=local)
        make_comm(cctx)
        g5 = vg.Reference(ref5, K=5, thread_num=meta5["index_thread_num"])
        cctx.set_reference(g5)
        lo, hi = shard_range(nck, rank, world)
        cctx.upload_reads(vg.PairedReads(*synth.pack_pairs(ck1[lo:hi], ck2[lo:hi])))
        md = cctx.map_pairs()
        cctx.pileup(False)
        cctx.pileup_allreduce()
        counts_all = cctx.counts()
        cons_all = cctx.consensus(ref5, True, 0)
        import hashlib
        digest = hashlib.sha256(md.records["cls"].tobytes() + b"".join(md.records["m"][f].tobytes() for f in PARITY_FIELDS[:-1])
                                + md.seq1_pool.tobytes()).hexdigest()
        gathered = [None] * world
        dist.all_gather_object(gathered, (digest, [int(x) for x in md.stats], hashlib.sha256(counts_all.tobytes() + cons_all.tobytes()).hexdigest()))
        if rank == 0:
            cctx.upload_reads(vg.PairedReads(*synth.pack_pairs(ck1, ck2)))
            full = cctx.map_pairs()
            cctx.pileup(False)   # no all-reduce: the single-GPU counts of the whole set
            ok_counts = hashlib.sha256(cctx.counts().tobytes() + cctx.consensus(ref5, True, 0).tobytes()).hexdigest()
            ok = all(g[2] == ok_counts for g in gathered)
            for r in range(world):
                a, b = shard_range(nck, r, world)
                sub = full.records[a:b]
                pres = (sub["m"]["present"] == 1)
                pool_lo = int(sub["m"]["seq1_offset"][pres].min()) if pres.any() else 0
                pool_hi = int((sub["m"]["seq1_offset"][pres] + sub["m"]["length"][pres]).max()) if pres.any() else 0
                d = hashlib.sha256(sub["cls"].tobytes() + b"".join(sub["m"][f].tobytes() for f in PARITY_FIELDS[:-1])
                                   + full.seq1_pool[pool_lo:pool_hi].tobytes()).hexdigest()
                ok = ok and d == gathered[r][0]
            ok = ok and (np.sum([g[1] for g in gathered], axis=0) == full.stats).all()
            multi_check = {"pairs": nck, "ranks": world, "ok": bool(ok),
                           "what": "a fixed config-5 set sharded over the ranks: every rank's records + sequence1 bytes == its block of "
                                   "the single-GPU result, summed statistics equal, NCCL-all-reduced counts and consensus on every rank "
                                   "== single-GPU counts and consensus"}
        cctx.close()

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
        # the ALGORITHMIC HBM bytes of the seeding stage are the read bases in (L1 + L2 per pair at 1 B/base)
        # and the candidate records out (12 B per window + 4 B per orientation); the LongHit stream between the two
        # kernels is an intermediate of this implementation, not algorithmic traffic. Both kernels are bound by
        # instruction issue on shared-memory data (issue_slot_frac below), so the HBM fraction is ~1e-3 by construction.
        algo_bytes = 2 * L * n + 12 * win_out + 4 * items
        stream_gen = 4 * L * n + 8 * lh_in                    # bytes the implementation actually moves: reads in both
        stream_win = 8 * lh_in + 12 * win_out + 4 * items     # orientations + the LongHit stream (8 B each)
        kern = {"seed_gen_kernel": (gen_ms, stream_gen), "seed_win_kernel": (win_ms, stream_win)}
        prof = {}
        pp = os.path.join(ROOT, "profiles", "kernel_metrics_r2.json")
        if os.path.exists(pp):
            prof = json.load(open(pp))
        rl = {}
        for k, (kms, kb) in kern.items():
            gbs = algo_bytes / (kms * 1e-3) / 1e9 if kms else None
            pk_ = prof.get("kernels", {}).get(k, {})
            rl[k] = {"bound": "hbm", "kernel": k, "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": gbs / hbm_peak if gbs else None,
                     "traffic": pk_["dram_bytes_per_pair"] * n if "dram_bytes_per_pair" in pk_ else None,
                     "traffic_source": ("static: ncu --set full capture of this kernel (%s), scaled to this launch" % prof.get("source", "profiles/"))
                     if "dram_bytes_per_pair" in pk_ else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes, "ms_per_step": kms,
                     "issue_slot_frac": pk_.get("issue_active_frac"), "warp_instr_per_orientation": pk_.get("warp_instr_per_orientation"),
                     "issue_source": "static: same ncu capture" if pk_ else None,
                     "stream_inclusive": {"bytes_per_launch": kb, "achieved": kb / (kms * 1e-3) / 1e9 if kms else None,
                                          "frac": kb / (kms * 1e-3) / 1e9 / hbm_peak if kms else None,
                                          "what": "with the LongHit stream between the two kernels counted (an intermediate, 8 B per LongHit)"},
                     "longhits_per_s": lh_in / (kms * 1e-3) if kms else None}
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
                    "ms_per_step": e2e_ms / args.steps, "over_device_resident": e2e_value / value,
                    "how": "%d host threads with one context each alternate over the steps (copies of one overlap kernels of the "
                           "other); each step = vg_upload_reads_packed (2 bits per base + exceptions, pinned) + vg_map_pairs + "
                           "vg_get_results_compact (records + run-length edit operations, from which sequence1 is a function of "
                           "the read) + vg_pileup + vg_consensus" % n_e2e},
            "e2e_byte_api": {"value": bytes_value, "unit": UNIT, "h2d_bytes_per_step": int(bytes_h2d), "d2h_bytes_per_step": int(bytes_d2h),
                             "ms_per_step": bytes_ms / args.steps,
                             "how": "the same with 1 byte per base up (vg_upload_reads) and the full sequence1 pool down (vg_get_results)"},
            "gpu_launches": int(launches2 - launches1),
            "clocks": clocks,
            "stage_ms_per_step": {"seed_pairselect": seed_ms, "seed_gen_kernel": gen_ms, "seed_win_kernel": win_ms,
                                  "seed_fallback_kernel": fb_ms, "sw_fill": fill_ms, "sw_traceback": tb_ms, "pileup": pu_ms,
                                  "map_total": tot_ms},
            "roofline": dict(rl[dominant], note="dominant kernel of the step. achieved = SURVEY 8(d) algorithmic bytes of the seeding "
                             "stage (read bases in + window records out) / this kernel's time: ~1e-3 of the HBM peak by construction — "
                             "the kernel is bound by instruction issue on shared-memory data (issue_slot_frac), not by HBM; "
                             "%d of %d read orientations took the general (fallback) kernel" % (int(fb_items), int(items))),
            "seed_gen_roofline": rl["seed_gen_kernel"],
            "seed_win_roofline": rl["seed_win_kernel"],
            "sw_roofline": {"bound": "dpx", "gcups_fill": cells / (fill_ms * 1e-3) / 1e9 if fill_ms else None,
                            "gcups_fill_traceback": cells / ((fill_ms + tb_ms) * 1e-3) / 1e9 if fill_ms else None,
                            "dpx_s16x2_ops_per_s_measured": dpx,
                            "gcups_peak_3ops_per_cellpair": dpx * 2 / 3 / 1e9,
                            "frac_of_3op_peak_fill": (cells / (fill_ms * 1e-3)) / (dpx * 2 / 3) if fill_ms else None,
                            # SASS of the inner loop (profiles/sass_r2.md): 6.6 ALU / DPX-pipe instructions per packed
                            # cell-pair (LOP3, 4 x VIADDMNMX, VIMNMX3, half a VIMNMX3 for the running maximum) + one IMAD.IADD
                            # on the FMA pipe; the measured peak is the issue rate of the ALU / DPX pipe
                            "alu_pipe_instr_per_cellpair_issued": 6.6,
                            "frac_of_dpx_issue_peak_fill": (cells / 2 * 6.6 / (fill_ms * 1e-3)) / dpx if fill_ms else None,
                            "cells_per_step": cells},
            "pileup_roofline": pile,
            "mapping_stats": [int(x) for x in stats],
        }
        if world > 1:
            out["multi_gpu_check"] = multi_check
            out["allreduce_us"] = float(np.median(allreduce_ms)) * 1e3 if allreduce_ms else None
            out["collective"] = "ncclAllReduce(int32, sum) of the pileup counts inside libvirgena_gpu.so (vg_pileup_allreduce); no host-language collective in the data path"
        threads = os.cpu_count() or 1
        if not args.no_msa:
            out["msa_mode"] = msa_arm(vg, synth, local, min(n, 200000))
        if not args.no_secondary:
            out["cfg4_fragmented"] = secondary_arm(vg, synth, local, 4, min(n, args.secondary_pairs), threads, 0 if args.no_cpu_baseline else 4000)
            out["cfg5_hcv_mixture"] = secondary_arm(vg, synth, local, 5, min(n, args.secondary_pairs), threads, 0 if args.no_cpu_baseline else 4000)
            out["random_model"] = random_model_arm(vg, synth, local, not args.no_cpu_baseline)
        if not args.no_cpu_baseline:
            ncpu = min(n, args.cpu_pairs or 3000 * threads)
            m1 = host_batches[0][0][:ncpu].numpy()
            m2 = host_batches[0][1][:ncpu].numpy()
            v, dt, res = cpu_baseline(ref, cut, m1, m2, ncpu, threads, meta["index_thread_num"], want_results=True)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                   "sample": "first %d pairs of batch 0 (%.1f s wall on %d threads, BatchSize=1000 tasks)" % (ncpu, dt, threads),
                                   "note": "C++ restatement of the reference (oracle/), not the JVM"}
            # the sample the oracle just mapped is the head of the batch the timed steps mapped: compare, outside the timed region
            bad = parity_check(ctx, ref, synth.pack_pairs(m1, m2), res)
            out["parity_checked_pairs"] = ncpu
            out["parity_ok"] = not bad
            if bad:
                out["parity_mismatch"] = bad
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
