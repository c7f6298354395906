"""Runs the SW fill+traceback kernels on N synthetic read-vs-window tasks (for ncu / timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import virgena_b200 as vg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L = int(sys.argv[3]) if len(sys.argv) > 3 else 250
W = int(sys.argv[4]) if len(sys.argv) > 4 else 372
ctx = vg.Context()
rng = np.random.default_rng(14)
al = np.frombuffer(b"ACGT", np.uint8)
w = al[rng.integers(0, 4, (n, W))]
o = (W - L) // 2
r = w[:, o:o + L].copy()
mut = rng.random(r.shape) < 0.1
r[mut] = al[rng.integers(0, 4, int(mut.sum()))]
s1o = np.arange(n + 1, dtype=np.int64) * L
s2o = np.arange(n + 1, dtype=np.int64) * W
for _ in range(reps):
    rec, pool, soff = ctx.sw_align_batch(r.reshape(-1), s1o, w.reshape(-1), s2o)
    t = ctx.timings()
    print("n=%d fill %.3f ms tb %.3f ms -> fill %.1f GCUPS, fill+tb %.1f GCUPS" % (
        n, t["sw_fill_ms"], t["traceback_ms"], t["sw_cells"] / t["sw_fill_ms"] / 1e6,
        t["sw_cells"] / (t["sw_fill_ms"] + t["traceback_ms"]) / 1e6))
sc = ctx.sw_max_score_batch(r.reshape(-1), s1o, w.reshape(-1), s2o)
t = ctx.timings()
print("score-only fill %.3f ms -> %.1f GCUPS" % (t["sw_fill_ms"], n * L * W / t["sw_fill_ms"] / 1e6))
print("dpx peak %.4e" % ctx.dpx_peak())
