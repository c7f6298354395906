"""Runs the mapping pipeline on N synthetic cfg3 pairs (for ncu / timing). args: n_pairs [reps] [cfg]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import virgena_b200 as vg
from virgena_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = int(sys.argv[3]) if len(sys.argv) > 3 else 3
d = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/cutoffs.json")))["cfg%d" % cfg]
cut = np.array(d["cutoffs"], np.int32)
ref, ends = synth.config_reference(cfg)
m1, m2 = synth.config_pairs(cfg, min(n, 20000))
reps_n = (n + len(m1) - 1) // len(m1)
m1 = np.tile(m1, (reps_n, 1))[:n]; m2 = np.tile(m2, (reps_n, 1))[:n]
ctx = vg.Context(K=5, cutoffs=cut)
ctx.set_reference(vg.Reference(ref, K=5, contig_ends=ends, fragment_ends=ends, is_fragmented=len(ends) > 1, thread_num=8))
ctx.upload_reads(vg.PairedReads(*synth.pack_pairs(m1, m2)))
for _ in range(reps):
    st = ctx.map_pairs(fetch=False)
    ctx.pileup()
    t = ctx.timings()
    print("n=%d seed %.2f ms fill %.2f tb %.2f pileup %.2f total %.2f -> %.0f pairs/s; stats %s" % (
        n, t["seed_ms"], t["sw_fill_ms"], t["traceback_ms"], t["pileup_ms"], t["total_ms"], n / t["total_ms"] * 1e3, st.tolist()))
