"""GPU parity: batched Smith-Waterman-Gotoh (fill + traceback) vs the oracle's restatement of
SmithWatermanGotoh.align / findMaxScore. Bit-exact on all 7 Alignment fields and sequence1."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _offs(seqs):
    return np.concatenate([[0], np.cumsum([len(x) for x in seqs])]).astype(np.int64)


def _check(ctx, o, s1, s2):
    rec, pool, soff = ctx.sw_align_batch(b"".join(s1), _offs(s1), b"".join(s2), _offs(s2))
    s1p, s2p = b"".join(s1), b"".join(s2)
    orec, opool, osoff = o.sw_batch(s1p if s1p else b"\0", _offs(s1), s2p if s2p else b"\0", _offs(s2), n_threads=8)
    bad = np.nonzero((rec != orec).any(axis=1))[0]
    assert len(bad) == 0, (bad[:5], rec[bad[:3]], orec[bad[:3]], [(s1[i], s2[i]) for i in bad[:1]])
    for i in range(len(s1)):
        n = rec[i][6]
        assert pool[soff[i]:soff[i] + n].tobytes() == opool[osoff[i]:osoff[i] + n].tobytes(), i
    sc = ctx.sw_max_score_batch(b"".join(s1), _offs(s1), b"".join(s2), _offs(s2))
    assert (sc == orec[:, 0]).all()
    return rec


@pytest.fixture(scope="module")
def env(built):
    import oracle
    import virgena_b200 as vg
    return vg.Context(), oracle.Oracle()


def test_reads_vs_windows(env):
    from virgena_b200 import synth
    ctx, o = env
    ref, _ = synth.config_reference(1)
    m1, m2 = synth.config_pairs(3, 1500)
    rng = np.random.default_rng(5)
    s1, s2 = [], []
    for i in range(1500):
        # a window near where the read came from is unknown here; use seeds from the oracle-free path:
        st = int(rng.integers(0, len(ref) - 380))
        n = int(rng.integers(250, 376))
        s1.append(m1[i].tobytes())
        s2.append(ref[st:st + n].tobytes())
    # half of the tasks get a true window: plant the read (mutated) inside
    for i in range(0, 1500, 2):
        st = int(rng.integers(0, len(ref) - 380))
        w = ref[st:st + 372].copy()
        r = w[60:310].copy()
        mut = rng.random(250) < 0.1
        r[mut] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(mut.sum()))]
        if i % 4 == 0:
            r = np.concatenate([r[:100], r[103:], r[:3]])  # a deletion in the read
        if i % 6 == 0:
            r = np.concatenate([r[:50], np.frombuffer(b"GGA", np.uint8), r[50:247]])  # an insertion
        s1[i] = r.tobytes()
        s2[i] = w.tobytes()
    rec = _check(ctx, o, s1, s2)
    assert rec[:, 0].max() > 200


def test_wide_checkpoints(env, monkeypatch):
    """The 8-byte checkpoint format (scores >= 4096 or gop - gep > 15) on the inputs that normally take the 4-byte one."""
    from virgena_b200 import synth
    monkeypatch.setenv("VG_SW_WIDE", "1")
    ctx, o = env
    ref, _ = synth.config_reference(1)
    m1, _ = synth.config_pairs(1, 300)
    rng = np.random.default_rng(8)
    s1, s2 = [], []
    for i in range(300):
        st = int(rng.integers(0, len(ref) - 380))
        w = ref[st:st + int(rng.integers(250, 376))].copy()
        r = w[10:10 + 240].copy() if i % 2 else m1[i]
        mut = rng.random(len(r)) < 0.12
        r = r.copy()
        r[mut] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(mut.sum()))]
        if i % 3 == 0:
            r = np.delete(r, slice(100, 104))
        s1.append(r.tobytes())
        s2.append(w.tobytes())
    _check(ctx, o, s1, s2)


def test_ragged_random(env):
    ctx, o = env
    rng = np.random.default_rng(11)
    al = np.frombuffer(b"ACGT", np.uint8)
    s1 = [al[rng.integers(0, 4, int(rng.integers(1, 300)))].tobytes() for _ in range(1200)]
    s2 = [al[rng.integers(0, 4, int(rng.integers(1, 420)))].tobytes() for _ in range(1200)]
    _check(ctx, o, s1, s2)


def test_short_and_long_reads(env):
    ctx, o = env
    rng = np.random.default_rng(12)
    al = np.frombuffer(b"ACGT", np.uint8)
    # 513 .. 1024 rows: contigs as queries (Graph.SWScore, ReferenceFinder.SWScoreCounter), 24 / 32 rows per lane
    for L, W in ((30, 45), (100, 150), (150, 225), (180, 200), (300, 450), (400, 520), (500, 600), (700, 760), (1024, 1100), (900, 400)):
        s2 = [al[rng.integers(0, 4, W)] for _ in range(101 if L <= 512 else 41)]
        s1 = []
        for w in s2:
            r = w[5:5 + L].copy() if L + 5 <= W else np.concatenate([al[rng.integers(0, 4, (L - W) // 2)], w, al[rng.integers(0, 4, L - W - (L - W) // 2)]])
            mut = rng.random(len(r)) < 0.08
            r[mut] = al[rng.integers(0, 4, int(mut.sum()))]
            s1.append(r.tobytes())
        _check(ctx, o, s1, [w.tobytes() for w in s2])


def test_sub_warp_bin_borders(env):
    """Read lengths on both sides of every border of the aligner's (lanes per pair, rows per lane) bins (8 x 8 up to 64
    bases, 16 x 6 up to 96, 16 x 8 up to 128, 32 x 6 up to 192, 32 x 8 up to 256, ...), odd task counts per bin (a group of
    lanes without a pair, a pair without a partner), empty reads and empty windows in between, reads planted in their windows
    so that the traceback walks every strip."""
    ctx, o = env
    rng = np.random.default_rng(21)
    al = np.frombuffer(b"ACGT", np.uint8)
    s1, s2 = [], []
    for L in (1, 7, 8, 9, 47, 48, 49, 63, 64, 65, 95, 96, 97, 127, 128, 129, 191, 192, 193, 255, 256, 257, 320, 321):
        for rep in range(3 + (L % 4)):                      # 3 .. 6 tasks of this length: odd and even bin populations
            W = int(L * rng.uniform(1.0, 1.5)) + int(rng.integers(0, 3))
            w = al[rng.integers(0, 4, W)]
            st = int(rng.integers(0, W - L + 1))
            r = w[st:st + L].copy()
            mut = rng.random(L) < 0.12
            r[mut] = al[rng.integers(0, 4, int(mut.sum()))]
            if L > 20 and rep % 2:
                r = np.concatenate([r[:L // 2], r[L // 2 + 2:], al[rng.integers(0, 4, 2)]])   # a deletion in the read
            s1.append(r.tobytes())
            s2.append(w.tobytes())
    s1 += [b"", b"ACGT" * 10, b"", b"A" * 64]
    s2 += [b"ACGTACGT", b"", b"", b"A" * 70]
    order = rng.permutation(len(s1))
    _check(ctx, o, [s1[i] for i in order], [s2[i] for i in order])
    for n in (1, 2, 3, 5):                                   # a handful of tasks: most groups of a warp stay empty
        _check(ctx, o, s1[:n], s2[:n])
        _check(ctx, o, s1[40:40 + n], s2[40:40 + n])


def test_short_reads_against_long_windows(env):
    """Reads of a sub-warp bin against windows too long for 32 / W of them in shared memory: the wider geometry takes over."""
    ctx, o = env
    rng = np.random.default_rng(22)
    al = np.frombuffer(b"ACGT", np.uint8)
    s1, s2 = [], []
    for L, W in ((20, 900), (60, 3000), (60, 5000), (100, 2500), (120, 6000), (50, 40)):
        for _ in range(3):
            w = al[rng.integers(0, 4, W)]
            st = int(rng.integers(0, max(1, W - L)))
            r = w[st:st + L].copy() if W >= L else al[rng.integers(0, 4, L)]
            mut = rng.random(len(r)) < 0.1
            r[mut] = al[rng.integers(0, 4, int(mut.sum()))]
            s1.append(r.tobytes())
            s2.append(w.tobytes())
    _check(ctx, o, s1, s2)


def test_edge_cases(env):
    ctx, o = env
    s1 = [b"AAAAAAAAAA", b"ACGTACGTAC", b"ACGTNNACGT", b"acgtACGTRYK", b"A", b"", b"ACGT", b"T" * 200, b"AC" * 100,
          b"ACGTACGTTTTTTTTACGATCGATCGA", b"GATTACA" * 20, b"NNNNNNNNNN", b"ACGTACGTAC" * 5]
    s2 = [b"CCCCCCCCCCCC", b"ACGTACGTAC", b"ACGTNNACGT", b"acgtACGTRYK", b"A", b"ACGT", b"", b"T" * 300, b"CA" * 150,
          b"ACGTACGTACGATCGATCGA", b"GATTACAGATTACA" * 8 + b"GG", b"NNNNNNNNNNNN", b"ACGTACGTAC" * 2 + b"GGG" + b"ACGTACGTAC" * 3]
    _check(ctx, o, s1, s2)


def test_other_scoring(env, built):
    import oracle
    import virgena_b200 as vg
    rng = np.random.default_rng(13)
    al = np.frombuffer(b"ACGT", np.uint8)
    s2 = [al[rng.integers(0, 4, 200)] for _ in range(300)]
    s1 = []
    for w in s2:
        r = w[20:160].copy()
        mut = rng.random(len(r)) < 0.15
        r[mut] = al[rng.integers(0, 4, int(mut.sum()))]
        r = np.concatenate([r[:40], r[44:]])
        s1.append(r.tobytes())
    for (ma, mi, go, ge) in ((1, -1, 1, 1), (5, -4, 10, 1), (2, -3, 5, 0), (3, -2, 2, 2)):
        ctx = vg.Context(match=ma, mismatch=mi, gop=go, gep=ge)
        o = oracle.Oracle(match=ma, mismatch=mi, gop=go, gep=ge)
        _check(ctx, o, s1, [w.tobytes() for w in s2])
        ctx.close()


_SW_FUZZ_SEEDS = list(range(8))
if __import__("os").environ.get("VG_FUZZ_SEEDS"):  # e.g. VG_FUZZ_SEEDS=8:100 for a longer one-off run
    _a, _b = __import__("os").environ["VG_FUZZ_SEEDS"].split(":")
    _SW_FUZZ_SEEDS = list(range(int(_a), int(_b)))


@pytest.mark.parametrize("seed", _SW_FUZZ_SEEDS)
def test_sw_fuzz(built, seed):
    """Differential fuzz of the aligner: random penalties (both checkpoint formats, gop < gep included), lengths 1 .. 1024
    on either side, related / unrelated / low-complexity sequences with N, '-', lower case and IUPAC bytes."""
    import oracle
    import virgena_b200 as vg
    rng = np.random.default_rng(3000 + seed)
    ma, mi = int(rng.integers(1, 6)), -int(rng.integers(1, 7))
    go, ge = int(rng.integers(1, 14)), int(rng.integers(0, 5))
    if seed % 4 == 3:
        go = ge + int(rng.integers(16, 40))  # wide checkpoints
    al = np.frombuffer(b"ACGT", np.uint8)
    odd = np.frombuffer(b"N-acgtRYKM", np.uint8)
    s1, s2 = [], []
    for _ in range(160):
        n = int(rng.choice([1, 5, 40, 150, 300, 700, 1024])) if rng.random() < 0.3 else int(rng.integers(1, 420))
        m = int(rng.choice([1, 7, 64, 250, 512, 800, 1024])) if rng.random() < 0.3 else int(rng.integers(1, 300))
        if ma * min(m, n) + ma + 1 > 32000:
            m = min(m, 300)
        w = al[rng.integers(0, 4, n)].copy()
        kind = rng.random()
        if kind < 0.6:      # the query is a mutated piece of the window (tiled when it is longer)
            r = np.resize(w[int(rng.integers(0, max(1, n - 1))):], m).copy()
            mut = rng.random(m) < rng.choice([0.0, 0.05, 0.25])
            r[mut] = al[rng.integers(0, 4, int(mut.sum()))]
            for _ in range(int(rng.integers(0, 3))):
                if len(r) > 12:
                    q = int(rng.integers(3, len(r) - 6))
                    r = np.delete(r, slice(q, q + int(rng.integers(1, 5)))) if rng.random() < 0.5 else \
                        np.insert(r, q, al[rng.integers(0, 4, int(rng.integers(1, 5)))])
            r = r[:1024]
        elif kind < 0.8:
            r = al[rng.integers(0, 4, m)]
        else:               # low complexity: many equal-score cells (tie-breaking)
            r = np.resize(al[rng.integers(0, 4, int(rng.integers(1, 4)))], m)
            w = np.resize(r[:int(rng.integers(1, 4))], n)
        r, w = r.copy(), w.copy()
        if rng.random() < 0.15:
            r[rng.integers(0, len(r), 2)] = odd[rng.integers(0, len(odd), 2)]
            w[rng.integers(0, len(w), 2)] = odd[rng.integers(0, len(odd), 2)]
        s1.append(r.tobytes())
        s2.append(w.tobytes())
    ctx = vg.Context(match=ma, mismatch=mi, gop=go, gep=ge)
    o = oracle.Oracle(match=ma, mismatch=mi, gop=go, gep=ge)
    _check(ctx, o, s1, s2)
    ctx.close()


def test_sw_throughput_and_dpx(env):
    ctx, o = env
    rng = np.random.default_rng(14)
    al = np.frombuffer(b"ACGT", np.uint8)
    n = 40000
    w = al[rng.integers(0, 4, (n, 372))]
    r = w[:, 60:310].copy()
    mut = rng.random(r.shape) < 0.1
    r[mut] = al[rng.integers(0, 4, int(mut.sum()))]
    s1o = np.arange(n + 1, dtype=np.int64) * 250
    s2o = np.arange(n + 1, dtype=np.int64) * 372
    rec, pool, soff = ctx.sw_align_batch(r.reshape(-1), s1o, w.reshape(-1), s2o)
    rec, pool, soff = ctx.sw_align_batch(r.reshape(-1), s1o, w.reshape(-1), s2o)
    t = ctx.timings()
    gcups_fill = t["sw_cells"] / (t["sw_fill_ms"] * 1e-3) / 1e9
    gcups_all = t["sw_cells"] / ((t["sw_fill_ms"] + t["traceback_ms"]) * 1e-3) / 1e9
    dpx = ctx.dpx_peak()
    print("\nSW fill %.3f ms, traceback %.3f ms -> %.1f GCUPS (fill), %.1f GCUPS (fill+tb); DPX peak %.3e s16x2 ops/s"
          % (t["sw_fill_ms"], t["traceback_ms"], gcups_fill, gcups_all, dpx))
    # spot-check parity on a sample
    for i in range(0, n, 997):
        rr, seq = o.sw_align(r[i].tobytes(), w[i].tobytes())
        assert (rec[i] == rr).all()
        assert pool[soff[i]:soff[i] + rr[6]].tobytes() == seq
