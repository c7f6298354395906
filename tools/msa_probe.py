"""MapperMSA.mapAllReads on N synthetic cfg2 pairs (timing). args: n_pairs [reps] [n_refs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import virgena_b200 as vg
from virgena_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n_refs = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = synth.config_msa(n_refs=n_refs)
ref, _ = synth.config_reference(1)
cut = oracle.cutoffs(ref, K=7, coef=1.5, read_num=800, min_len=250, max_len=250, step=10, seed=5, msa_rows=rows)
m1, m2 = synth.config_pairs(2, min(n, 5000))
k = (n + len(m1) - 1) // len(m1)
m1 = np.tile(m1, (k, 1))[:n]; m2 = np.tile(m2, (k, 1))[:n]
aln = vg.ReferenceAlignment(rows, K=7)
mapper = vg.MapperMSA(K=7, indel_tolerance=1.5, cutoffs=cut)
reads = vg.PairedReads(*synth.pack_pairs(m1, m2))
for _ in range(reps):
    t0 = time.time()
    md = mapper.mapAllReads(reads, aln)
    dt = time.time() - t0
    t = mapper.ctx.timings()
    print("msa n=%d refs=%d columns=%d seed %.2f ms wall %.1f ms -> %.0f pairs/s; stats %s" % (n, n_refs, rows.shape[1], t["seed_ms"], dt * 1e3, n / t["seed_ms"] * 1e3, md.stats.tolist()))
