"""Multi-GPU through the C ABI: the sharded path must reproduce the single-GPU result bit for bit —
gathered records, sequence1 bytes, statistics, all-reduced pileup counts and consensus
(ConsensusBuilderSimple.java:138-145). Needs >= 2 GPUs (run with `gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import virgena_b200 as vg
    from virgena_b200 import _lib
    return _lib.lib().vg_device_count()


def _cfg5(n):
    import json
    from virgena_b200 import synth
    ref, ends = synth.config_reference(5)
    cut = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "cutoffs.json")))["cfg5"]["cutoffs"], np.int32)
    m1, m2 = synth.config_pairs(5, n)
    return ref, cut, synth.pack_pairs(m1, m2)


@pytest.mark.parametrize("n_dev", [2, 3, 4, 8])
def test_vg_multi_equals_single_gpu(built, n_dev):
    """One process, n_dev devices (vg_multi_*, ncclCommInitAll): a fixed config-5 set (six-genotype HCV mixture mapped to
    genotype 1) sharded over the devices == the same set on one device."""
    import virgena_b200 as vg
    if _n_gpus() < n_dev:
        pytest.skip("needs %d GPUs" % n_dev)
    n = 6001  # not divisible: shards differ by one pair
    ref, cut, packed = _cfg5(n)
    reads = vg.PairedReads(*packed)
    genome = vg.Reference(ref, K=5, thread_num=8)
    single = vg.Mapper(K=5, cutoffs=cut, device=0)
    md1 = single.mapReads(reads, genome)
    cons1 = vg.ConsensusBuilderSimple(single).getConsensusSimple(True, 0, genome.seqB)
    counts1 = single.ctx.counts()
    multi = vg.Mapper(K=5, cutoffs=cut, devices=list(range(n_dev)))
    mdN = multi.mapReads(reads, genome)
    assert mdN.records.tobytes() == md1.records.tobytes()
    assert mdN.seq1_pool.tobytes() == md1.seq1_pool.tobytes()
    assert (mdN.stats == md1.stats).all()
    consN = vg.ConsensusBuilderSimple(multi).getConsensusSimple(True, 0, genome.seqB)
    assert (multi.ctx.counts() == counts1).all() and counts1.sum() > 0
    assert consN.tobytes() == cons1.tobytes()
    # every device holds the global counts after the all-reduce
    for d in range(n_dev):
        assert (multi.ctx.device_ctx(d).counts() == counts1).all()
    # a class mix very different from config 3: most pairs come from the five other genotypes
    assert len(md1.unmapped) + len(md1.discordant) + len(md1.leftMateMapped) + len(md1.rightMateMapped) > 0
    # remap of a subset (Mapper.java:477): results come back in the order of the index list, whichever device owns a pair
    sub = np.concatenate([np.arange(n - 1, 0, -7), np.arange(3, n, 11)]).astype(np.int64)
    s1 = single.ctx.map_pairs(sub)
    sN = multi.ctx.map_pairs(sub)
    assert sN.records.tobytes() == s1.records.tobytes() and sN.seq1_pool.tobytes() == s1.seq1_pool.tobytes()
    assert (sN.stats == s1.stats).all()
    multi.ctx.pileup()
    single.ctx.pileup()
    assert (multi.ctx.counts() == single.ctx.counts()).all()
    # packed upload + compact download through the multi layer == the single device's (records and operations byte for byte)
    reads = vg.PairedReads(*packed)
    multi.ctx.upload_reads_packed(reads)
    single.ctx.upload_reads_packed(reads)
    for sub_idx in (None, sub):
        a1 = single.ctx.map_pairs(sub_idx, fetch=False)
        aN = multi.ctx.map_pairs(sub_idx, fetch=False)
        assert (np.asarray(a1) == np.asarray(aN)).all()
        r1, o1 = single.ctx.get_results_compact()
        rN, oN = multi.ctx.get_results_compact()
        assert rN.tobytes() == r1.tobytes() and oN.tobytes() == o1.tobytes()
    multi.ctx.close()
    single.ctx.close()


_RANK_SCRIPT = r"""
import os, sys, json, time
import numpy as np
sys.path.insert(0, sys.argv[1])
rank, world, n, tmp = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
import virgena_b200 as vg
from virgena_b200 import synth
from virgena_b200.dist import shard_range
ref, ends = synth.config_reference(5)
cut = np.array(json.load(open(os.path.join(sys.argv[1], "tests", "golden", "cutoffs.json")))["cfg5"]["cutoffs"], np.int32)
m1, m2 = synth.config_pairs(5, n)
lo, hi = shard_range(n, rank, world)
ctx = vg.Context(K=5, cutoffs=cut, device=rank)
idf = os.path.join(tmp, "nccl_id.bin")
if rank == 0:
    uid = vg.Context.comm_unique_id()
    with open(idf + ".tmp", "wb") as f:
        f.write(uid.tobytes())
    os.replace(idf + ".tmp", idf)          # the id travels by a file: no host-language collective anywhere
else:
    t0 = time.time()
    while not os.path.exists(idf):
        if time.time() - t0 > 120:
            raise SystemExit("no id file")
        time.sleep(0.05)
    uid = np.fromfile(idf, np.uint8)
ctx.comm_init_rank(uid, rank, world)
genome = vg.Reference(ref, K=5, thread_num=8)
ctx.set_reference(genome)
ctx.upload_reads(vg.PairedReads(*synth.pack_pairs(m1[lo:hi], m2[lo:hi])))
md = ctx.map_pairs()
ctx.pileup(False)
ctx.pileup_allreduce()
cons = ctx.consensus(ref, True, 0)
np.savez(os.path.join(tmp, "rank%d.npz" % rank), rec=md.records, pool=md.seq1_pool, stats=md.stats, counts=ctx.counts(), cons=cons,
         ms=ctx.allreduce_ms())
"""


def test_one_process_per_gpu_equals_single_gpu(built, tmp_path):
    """One process per GPU (vg_comm_unique_id / vg_comm_init_rank / vg_pileup_allreduce): the way bench.py --gpus N runs under
    torchrun, here with plain subprocesses and the NCCL id handed over in a file."""
    import virgena_b200 as vg
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    world, n = 2, 4001
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), str(world), str(n), str(tmp_path)]) for r in range(world)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    ref, cut, packed = _cfg5(n)
    single = vg.Mapper(K=5, cutoffs=cut, device=0)
    genome = vg.Reference(ref, K=5, thread_num=8)
    md1 = single.mapReads(vg.PairedReads(*packed), genome)
    cons1 = vg.ConsensusBuilderSimple(single).getConsensusSimple(True, 0, genome.seqB)
    parts = [np.load(tmp_path / ("rank%d.npz" % r)) for r in range(world)]
    rec = np.concatenate([p["rec"] for p in parts])
    base = 0
    for p in parts:  # sequence1 offsets are local to a rank's pool
        pass
    pools = [p["pool"] for p in parts]
    off = np.cumsum([0] + [len(x) for x in pools])
    at = 0
    for r, p in enumerate(parts):
        k = len(p["rec"])
        sl = rec[at:at + k]
        pres = (sl["m"]["present"] == 1) & (sl["m"]["seq1_offset"] >= 0)
        sl["m"]["seq1_offset"][pres] += off[r]
        at += k
    assert rec.tobytes() == md1.records.tobytes()
    assert np.concatenate(pools).tobytes() == md1.seq1_pool.tobytes()
    assert (sum(p["stats"] for p in parts) == md1.stats).all()
    for p in parts:
        assert (p["counts"] == single.ctx.counts()).all()
        assert p["cons"].tobytes() == cons1.tobytes()
