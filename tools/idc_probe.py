"""Debug probe: identity counting GPU vs oracle on a small batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, oracle, virgena_b200 as vg
from virgena_b200 import synth
ref, _ = synth.config_reference(1)
cut = np.full(600, 30, np.int32)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
window, threads = 50, 1
m1, m2 = synth.config_pairs(3, n)
rng = np.random.default_rng(5001)
r1 = [m1[i].tobytes()[:int(rng.integers(20, 251))] for i in range(n)]
r2 = [m2[i].tobytes()[int(rng.integers(0, 30)):] for i in range(n)]
bases, off1, len1, off2, len2 = synth.pack_ragged(r1, r2)
o = oracle.Oracle(K=5, cutoffs=cut)
o.set_reference(ref, index_thread_num=8)
orec, opool, _ = o.map_pairs(bases, off1, len1, off2, len2, n_threads=8)
mapper = vg.Mapper(K=5, cutoffs=cut)
md = mapper.mapReads(vg.PairedReads(bases, off1, len1, off2, len2), vg.Reference(ref, K=5, thread_num=8))
for k in range(0):
    sub = orec[k:k + 1].copy()
    lo = min([int(sub["m"][0][j]["seq1_offset"]) for j in range(2) if sub["m"][0][j]["present"]] + [1 << 60])
    wi, wc = oracle.identity_counting(sub, opool, len1[k:k + 1], len2[k:k + 1], ref, window, threads)
    print(k, [(int(sub["m"][0][j]["present"]), int(sub["m"][0][j]["start"]), int(sub["m"][0][j]["start1"]), int(sub["m"][0][j]["end1"]),
               int(sub["m"][0][j]["start2"]), int(sub["m"][0][j]["end2"]), int(sub["m"][0][j]["length"])) for j in range(2)], int(wc.sum()))
want_i, want_c = oracle.identity_counting(orec, opool, len1, len2, ref, window, threads)
got_i, got_c = mapper.ctx.identity_counting(window, threads)
rec = md.records
for f in ("present", "start", "start1", "end1", "start2", "end2", "length", "identity"):
    d = np.nonzero(rec["m"][f] != orec["m"][f])
    print("records differ in", f, len(d[0]))
gi2, gc2 = oracle.identity_counting(rec.astype(oracle.PAIR_DTYPE), md.seq1_pool, len1, len2, ref, window, threads)
print("oracle on GPU records vs oracle:", int((gc2 != want_c).sum()), " vs gpu:", int((gc2 != got_c).sum()))
cum = 0
for k in range(n):
    sub = rec[k:k + 1].astype(oracle.PAIR_DTYPE)
    wi, wc = oracle.identity_counting(sub, md.seq1_pool, len1[k:k + 1], len2[k:k + 1], ref, window, threads)
    if wc[4927:5128].all() and wc[4926] == 0 and wc[5128] == 0:
        print("candidate", k, sub["m"][0], len1[k], len2[k])
bad = np.nonzero(got_c != want_c)[0]
print("cov mismatches", len(bad), bad[:20], got_c[bad[:20]], want_c[bad[:20]], "sum", got_c.sum(), want_c.sum())
badi = np.nonzero(got_i.view(np.uint32) != want_i.view(np.uint32))[0]
print("identity mismatches", len(badi), badi[:10], got_i[badi[:10]], want_i[badi[:10]])

if len(bad):
    x = int(bad[0])
    for k in range(n):
        sub = orec[k:k + 1].copy()
        wi, wc = oracle.identity_counting(sub, opool, len1[k:k + 1], len2[k:k + 1], ref, window, threads)
        if wc[x]:
            mp = vg.Mapper(K=5, cutoffs=cut)
            b1, o1, l1, o2, l2 = synth.pack_ragged([r1[k]], [r2[k]])
            mp.mapReads(vg.PairedReads(b1, o1, l1, o2, l2), vg.Reference(ref, K=5, thread_num=8))
            gi, gc = mp.ctx.identity_counting(window, threads)
            mp.ctx.close()
            if (gc == wc).all():
                continue
            print("DIFF", np.nonzero(gc != wc)[0][[0, -1]])
            print("covering", k, [(int(sub["m"][0][j]["present"]), int(sub["m"][0][j]["start"]), int(sub["m"][0][j]["start1"]), int(sub["m"][0][j]["end1"]),
               int(sub["m"][0][j]["start2"]), int(sub["m"][0][j]["end2"]), int(sub["m"][0][j]["length"])) for j in range(2)], len1[k], len2[k], int(wc[x]))
