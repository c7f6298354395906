"""Generates tests/golden/small_case.npz with the oracle: a small seeded mapping case (300 pairs of config 1 with
ragged / absent / N-containing reads) and all of its outputs. The reference itself cannot run here (no JVM), so
this freezes the oracle's answers (which the hand-derived KATs pin).  Usage: python tests/golden/make_golden.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

import oracle
from virgena_b200 import synth

here = os.path.dirname(os.path.abspath(__file__))
ref, ends = synth.config_reference(1)
cut = np.array(json.load(open(os.path.join(here, "cutoffs.json")))["cfg1"]["cutoffs"], np.int32)
m1, m2 = synth.config_pairs(1, 240)
d1, d2 = synth.config_pairs(3, 60)
rng = np.random.default_rng(42)
r1 = [x.tobytes() for x in m1] + [x.tobytes() for x in d1]
r2 = [x.tobytes() for x in m2] + [x.tobytes() for x in d2]
for i in range(0, 300, 13):
    r1[i] = r1[i][:int(rng.integers(40, 250))]
for i in range(5, 300, 37):
    r2[i] = b"*"
for i in range(7, 300, 41):
    r1[i] = r1[i][:100] + b"NNN" + r1[i][103:]
bases, off1, len1, off2, len2 = synth.pack_ragged(r1, r2)
o = oracle.Oracle(K=5, cutoffs=cut)
o.set_reference(ref, index_thread_num=8)
rec, pool, st = o.map_pairs(bases, off1, len1, off2, len2)
counts = oracle.pileup(rec, pool, len(ref))
cons = oracle.consensus(counts, ref, True, 0)
np.savez_compressed(os.path.join(here, "small_case.npz"), ref=ref, cutoffs=cut, index_thread_num=8, bases=bases, off1=off1,
                    len1=len1, off2=off2, len2=len2, records=rec, seq1_pool=pool, stats=st, counts=counts, consensus=cons)
print("classes", np.bincount(rec["cls"], minlength=6), "stats", st)
