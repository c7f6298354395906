"""Smith-Waterman fill + traceback GCUPS by read length (uniform batches, windows 1.0..1.5 x the read).
args: n_tasks lengths(comma separated)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import virgena_b200 as vg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
lengths = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "40,60,90,120,150,180,210,250").split(",")]
rng = np.random.default_rng(15)
al = np.frombuffer(b"ACGT", np.uint8)
ctx = vg.Context()
for L in lengths:
    lens = np.full(n, L, np.int64)
    wl = (lens * rng.uniform(1.0, 1.5, n)).astype(np.int64)
    s1o = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    s2o = np.concatenate([[0], np.cumsum(wl)]).astype(np.int64)
    w = al[rng.integers(0, 4, int(s2o[-1]))]
    # the read is a mutated slice of its window (10 % substitutions): the traceback walks the whole read
    o = (wl - lens) // 2
    idx = (s2o[:-1] + o)[:, None] + np.arange(L)[None, :]
    r = w[idx].reshape(-1).copy()
    mut = rng.random(len(r)) < 0.1
    r[mut] = al[rng.integers(0, 4, int(mut.sum()))]
    cells = float((lens * wl).sum())
    for _ in range(2):
        ctx.sw_align_batch(r, s1o, w, s2o)
        t = ctx.timings()
    print("L=%3d n=%d cells %.3e: fill %.3f ms tb %.3f ms -> fill %.1f GCUPS, fill+tb %.1f GCUPS %s" % (
        L, n, cells, t["sw_fill_ms"], t["traceback_ms"], cells / t["sw_fill_ms"] / 1e6,
        cells / (t["sw_fill_ms"] + t["traceback_ms"]) / 1e6, os.environ.get("VG_SW_MINR", "")))
