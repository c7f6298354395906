"""Generates tests/golden/cutoffs.json: KMerCounterBase.cutoffs for the synthetic configs.

The reference derives `cutoffs` from a random model driven by an UNSEEDED java.util.Random
(MarkovModel.java:94), so they are an *input* of the hot path. This script runs the
oracle's seeded restatement of RandomModel.genModel + computeCuttofs (config.xml defaults: Order 4,
ReadNum 10000, Step 10, pValue 0.01) once and freezes the result, so that bench.py's GPU arm does not
need the oracle at run time.   Usage: python tests/golden/make_cutoffs.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

import oracle
from virgena_b200 import synth

out = {}
for cfg, L in ((1, 250), (4, 150), (5, 150)):
    ref, ends = synth.config_reference(cfg)
    cut = oracle.cutoffs(ref, K=5, coef=1.25, order=4, read_num=10000, min_len=L, max_len=L, step=10, p_value=0.01,
                         contig_ends=ends, is_fragmented=len(ends) > 1, index_thread_num=8, seed=synth.SEED0 + cfg)
    out["cfg%d" % cfg] = {"K": 5, "L": L, "index_thread_num": 8, "cutoffs": cut.tolist()}
    print(cfg, cut[L])
rows = synth.config_msa()
ref, _ = synth.config_reference(1)
cut = oracle.cutoffs(ref, K=7, coef=1.5, order=4, read_num=10000, min_len=250, max_len=250, step=10, p_value=0.01,
                     seed=synth.SEED0 + 2, msa_rows=rows)
out["cfg2"] = {"K": 7, "L": 250, "cutoffs": cut.tolist()}
print(2, cut[250])
out["cfg3"] = out["cfg1"]
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cutoffs.json"), "w") as f:
    json.dump(out, f)
