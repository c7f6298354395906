"""ContigBuilder.countReads throughput: n_clusters centroids of ~600 bases, reads_per_cluster reads of 150-250
bases each. args: n_clusters [reads_per_cluster]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import virgena_b200 as vg
from virgena_b200 import synth
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rp = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ref, _ = synth.config_reference(1)
rng = np.random.default_rng(1)
cents, reads, rc = [], [], []
for c in range(nc):
    ln = int(rng.integers(400, 800))
    p = int(rng.integers(0, len(ref) - ln))
    cen = ref[p:p + ln]
    cents.append(cen.tobytes())
    for _ in range(rp):
        L = int(rng.integers(150, 251))
        q = int(rng.integers(0, ln - L))
        r = cen[q:q + L].copy()
        m = rng.random(L) < 0.05
        r[m] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(m.sum()))]
        reads.append(r.tobytes())
        rc.append(c)
mapper = vg.Mapper(K=5, cutoffs=np.full(400, 40, np.int32))
cb = vg.ContigBuilder(mapper)
cpool = np.frombuffer(b"".join(cents), np.uint8)
coff = np.zeros(nc + 1, np.int64); np.cumsum([len(c) for c in cents], out=coff[1:])
rpool = np.frombuffer(b"".join(reads), np.uint8)
roff = np.zeros(len(reads) + 1, np.int64); np.cumsum([len(r) for r in reads], out=roff[1:])
rcv = np.array(rc, np.int32)
for rep in range(3):
    t0 = time.time()
    win, found = mapper.ctx.map_to_regions(cpool, coff, rpool, roff, rcv)
    t1 = time.time()
    print("map_to_regions: %d reads against %d centroids: %.1f ms (%.0f reads/s, host table build + copies included); found %d" % (
        len(reads), nc, (t1 - t0) * 1e3, len(reads) / (t1 - t0), int((found != 0).sum())))
t0 = time.time()
out = cb.countReads(cents, reads, rc)
print("countReads (seed + align + pileup, python glue included): %.1f ms" % ((time.time() - t0) * 1e3))
