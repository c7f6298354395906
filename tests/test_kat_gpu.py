"""The hand-derived known-answer vectors of tests/test_oracle_kat.py run against the CUDA path itself, through the
C ABI: here the device is compared with values worked out by hand from the Java source (derivations and file:line
citations are in test_oracle_kat.py), not with the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _offs(seqs):
    return np.concatenate([[0], np.cumsum([len(x) for x in seqs])]).astype(np.int64)


# (reference, K, index threads, contig ends, read, LongHits after mergeHits + splitLongHits)
HIT_CASES = [
    (b"TTACGATCGGAA", 3, 1, None, b"ACGATCGG", [[2, 0, 5]]),                       # one diagonal chain
    (b"ACGTTACGA", 3, 1, None, b"ACGA", [[0, 0, 0], [5, 0, 1]]),
    (b"AAAAAA", 3, 1, None, b"AAAA", [[0, 0, 1]]),                                 # shorter overlapping hits trimmed away
    (b"ACGTACGTTT", 3, 1, None, b"CGTACGTT", [[1, 0, 5]]),                         # longer later hit truncates prev
    (b"ACGTACGT", 3, 2, None, b"TACGT", [[0, 1, 0], [3, 0, 2]]),                   # Q1 duplicate adds a zero-length hit
    (b"ACGTACGT", 3, 1, None, b"TACGT", [[0, 1, 0], [3, 0, 2]]),
    (b"AACAGATCCGCTGGTA", 3, 2, None, b"CAGATCTCCGCT", [[2, 0, 3], [8, 8, 0]]),    # Q2 key collision on re-insert
    (b"AACAGATCCGCTGGTA", 3, 1, None, b"CAGATCTCCGCT", [[2, 0, 3], [8, 8, 1]]),
    (b"ACGGTCATTGCATTTTTTTTTTTTTTTTTTTTACGGTCATTGCA", 4, 1, None, b"ACGGTCATTGCA", [[0, 0, 8], [32, 0, 8]]),
    (b"ACGGTCATTGCAAGCT", 4, 1, None, b"ACGGTCATGTGCAAGCT", [[0, 0, 4], [8, 9, 4]]),
]


@pytest.mark.parametrize("case", range(len(HIT_CASES)))
def test_long_hits_known_answers(built, case):
    import virgena_b200 as vg
    ref, K, threads, ends, read, want = HIT_CASES[case]
    ctx = vg.Context(K=K, cutoffs=np.zeros(64, np.int32))
    ctx.set_reference(vg.Reference(ref, K=K, contig_ends=ends, thread_num=threads))
    assert ctx.debug_seed_hits(read).tolist() == want
    ctx.close()


def test_split_long_hits_known_answer(built):
    import virgena_b200 as vg
    ref = b"ACGTTGCAGGATCCAT"
    ctx = vg.Context(K=3, cutoffs=np.zeros(64, np.int32))
    ctx.set_reference(vg.Reference(ref, K=3, contig_ends=[8, 16], thread_num=1))
    h = ctx.debug_seed_hits(ref[3:13]).tolist()
    assert [3, 0, 2] in h and [8, 5, 2] in h
    assert all(not (g < 8 < g + 3 + ln) for g, r, ln in h)
    ctx.close()


def test_windows_known_answers(built):
    import virgena_b200 as vg
    W = b"ACGGTCATTGCA"
    ref = W + b"TTTTTTTTTTTTTTTTTTTT" + W
    cut = np.zeros(64, np.int32)
    ctx = vg.Context(K=4, cutoffs=cut)
    ctx.set_reference(vg.Reference(ref, K=4, thread_num=1))
    win, nwin = ctx.seed_batch(W, _offs([W]), max_windows=8)
    assert nwin[0] == 2 and win[0, :2].tolist() == [[0, 12, 8], [32, 44, 8]]
    ctx.close()
    cut[12] = 8  # strict: count > cutoffs[L]
    ctx = vg.Context(K=4, cutoffs=cut)
    ctx.set_reference(vg.Reference(ref, K=4, thread_num=1))
    win, nwin = ctx.seed_batch(W, _offs([W]), max_windows=8)
    assert nwin[0] == 0
    # touching windows (end == start) do not overlap
    ctx2 = vg.Context(K=4, cutoffs=np.zeros(64, np.int32))
    ctx2.set_reference(vg.Reference(W + W, K=4, thread_num=1))
    win, nwin = ctx2.seed_batch(W, _offs([W]), max_windows=8)
    assert nwin[0] == 2 and win[0, :2].tolist() == [[0, 12, 8], [12, 24, 8]]
    ctx.close()
    ctx2.close()


def test_concordance_bonus_q3_known_answer(built):
    """Q3: lowCoef = 0 (single-reference Mapper) gives the adjacency bonus, lowCoef = 1/1.05 does not."""
    import virgena_b200 as vg
    ref = b"ACGGTCATTGCAAGCT"
    read = b"ACGGTCATGTGCAAGCT"
    for lo, hi, count in ((0.0, 1.25, 9), (1 / 1.05, 1.05, 8)):
        ctx = vg.Context(K=4, low_coef=lo, high_coef=hi, cutoffs=np.zeros(64, np.int32))
        ctx.set_reference(vg.Reference(ref, K=4, thread_num=1))
        win, found = ctx.best_window_batch(read, _offs([read]))
        assert found[0] == 1 and win[0].tolist() == [0, 16, count]
        ctx.close()


def test_map_read_to_region_known_answers(built):
    import virgena_b200 as vg
    ctx = vg.Context(K=5, cutoffs=np.zeros(64, np.int32))
    ctx.set_reference(vg.Reference(b"AAAAACCCCCGGGGGTTTTT", K=5, thread_num=1))
    reads = [b"CCCCCGGGGG", b"CCCCCTTTTT", b"ACGTACGTAC"]
    win, found = ctx.best_window_batch(b"".join(reads), _offs(reads))
    assert found.tolist() == [1, 1, 0]
    assert win[0].tolist() == [5, 15, 5] and win[1].tolist() == [5, 17, 0]
    ctx.close()


SW_CASES = [
    # (match, mismatch, gop, gep), s1, s2, [score, start1, end1, start2, end2, identity, length], sequence1
    ((2, -3, 5, 2), b"ACGTACGT", b"ACGTACGT", [16, 0, 8, 0, 8, 8, 8], b"ACGTACGT"),
    ((2, -3, 5, 2), b"AAAA", b"CCCC", [0, 0, 0, 0, 0, 0, 0], b""),                                   # Q8
    ((2, -3, 5, 2), b"ACGTTACGT", b"ACGTAACGT", [13, 0, 9, 0, 9, 8, 9], b"ACGTTACGT"),
    ((2, -3, 5, 2), b"ACGGTCATTGCAAGCTGGAT", b"ACGGTCATTGTTCAAGCTGGAT", [33, 0, 20, 0, 22, 20, 22], b"ACGGTCATTG--CAAGCTGGAT"),
    ((2, -3, 5, 2), b"ACGGTCATTGTTTCAAGCTGGAT", b"ACGGTCATTGCAAGCTGGAT", [31, 0, 23, 0, 20, 20, 23], b"ACGGTCATTGtttCAAGCTGGAT"),
    ((2, -3, 5, 2), b"ACGT", b"ACGTNNACGT", [8, 0, 4, 0, 4, 4, 4], b"ACGT"),                         # first maximum, row-major
    ((2, -3, 5, 2), b"ACGTTTACGT", b"ACGT", [8, 0, 4, 0, 4, 4, 4], b"ACGT"),
    ((2, -3, 5, 2), b"ACGGTCATTGRCAAGCTGGAT", b"ACGGTCATTGCAAGCTGGAT", None, b"ACGGTCATTG\x00CAAGCTGGAT"),  # Q9
    ((2, -1, 3, 1), b"AAGCA", b"AACAGGA", [5, 1, 5, 0, 4, 3, 4], b"AGCA"),                            # v == f == g: DIAGONAL
    ((2, -3, 5, 2), b"CCCAGCAA", b"CCCGAGCA", [9, 0, 7, 1, 8, 6, 7], b"CCCAGCA"),                     # v == f == h: DIAGONAL
    ((1, -1, 1, 1), b"CACAGA", b"ACAAGAA", [4, 0, 6, 1, 6, 5, 6], b"CAcAGA"),                         # v == g == h: UP
]


@pytest.mark.parametrize("case", range(len(SW_CASES)))
def test_smith_waterman_known_answers(built, case):
    import virgena_b200 as vg
    (ma, mi, go, ge), s1, s2, want, seq = SW_CASES[case]
    ctx = vg.Context(K=5, cutoffs=np.zeros(64, np.int32), match=ma, mismatch=mi, gop=go, gep=ge)
    # the same task several times and next to a different partner: both 16-bit lanes of the packed kernel
    s1s = [s1, s1, b"ACGTACGTAA", s1]
    s2s = [s2, s2, b"TTACGTACGG", s2]
    rec, pool, soff = ctx.sw_align_batch(b"".join(s1s), _offs(s1s), b"".join(s2s), _offs(s2s))
    for i in (0, 1, 3):
        if want is not None:
            assert rec[i].tolist() == want, (i, rec[i])
        assert pool[soff[i]:soff[i] + rec[i, 6]].tobytes() == seq
    sc = ctx.sw_max_score_batch(b"".join(s1s), _offs(s1s), b"".join(s2s), _offs(s2s))
    assert sc[0] == rec[0, 0] and sc[3] == rec[3, 0]
    ctx.close()


def test_pileup_and_consensus_known_answers(built):
    """ConsensusBuilderSimple.java:56-75 (lower case skipped, '-' -> column 4, N -> A: Q11) and :98-120 (first maximum
    wins, <= threshold -> genome / gap) through vg_pileup_alignments + a mapping-free consensus check."""
    import virgena_b200 as vg
    ctx = vg.Context(K=5, cutoffs=np.zeros(64, np.int32))
    seq1 = b"ANt-G"
    counts = ctx.pileup_alignments(seq1, [0], [len(seq1)], [3], 10)
    exp = np.zeros((10, 5), np.int32)
    exp[3, 0] = 1
    exp[4, 0] = 1
    exp[5, 4] = 1
    exp[6, 2] = 1
    assert (counts == exp).all()
    ctx.close()
