"""Smith-Waterman fill + traceback on a ragged batch (read lengths 50..250, windows 1.0..1.5 x the read) against a uniform
batch of the same number of cells: GCUPS of both, with and without length binning (VG_SW_NOBIN=1).  args: n_tasks [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import virgena_b200 as vg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rng = np.random.default_rng(15)
al = np.frombuffer(b"ACGT", np.uint8)


def batch(lens):
    wl = (lens * rng.uniform(1.0, 1.5, len(lens))).astype(np.int64)
    s1o = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    s2o = np.concatenate([[0], np.cumsum(wl)]).astype(np.int64)
    w = al[rng.integers(0, 4, int(s2o[-1]))]
    r = np.empty(int(s1o[-1]), np.uint8)
    for i in range(len(lens)):  # the read is a mutated slice of its window
        o = (wl[i] - lens[i]) // 2
        r[s1o[i]:s1o[i + 1]] = w[s2o[i] + o:s2o[i] + o + lens[i]]
    mut = rng.random(len(r)) < 0.1
    r[mut] = al[rng.integers(0, 4, int(mut.sum()))]
    return r, s1o, w, s2o, float((lens * wl).sum())


ctx = vg.Context()
for name, lens in (("uniform 250", np.full(n, 250, np.int64)), ("ragged 50..250", rng.integers(50, 251, n).astype(np.int64))):
    r, s1o, w, s2o, cells = batch(lens)
    for _ in range(reps):
        ctx.sw_align_batch(r, s1o, w, s2o)
        t = ctx.timings()
    print("%-16s n=%d cells %.3e: fill %.3f ms tb %.3f ms -> fill %.1f GCUPS, fill+tb %.1f GCUPS%s" % (
        name, n, cells, t["sw_fill_ms"], t["traceback_ms"], cells / t["sw_fill_ms"] / 1e6,
        cells / (t["sw_fill_ms"] + t["traceback_ms"]) / 1e6, " (VG_SW_NOBIN)" if os.environ.get("VG_SW_NOBIN") else ""))
