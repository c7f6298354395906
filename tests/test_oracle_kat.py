"""Known-answer tests that pin the oracle (oracle/oracle.cpp).

The reference ships no tests, fixtures or golden vectors and cannot run here, so each case
below was worked out BY HAND from the Java source (file:line cited) — one per quirk of SURVEY.md §8a.
They run on the CPU (no GPU needed)."""
import numpy as np
import pytest

import oracle


def _o(**kw):
    return oracle.Oracle(**kw)


# --------------------------------------------------------------------------------------------------
# Constants.java
# --------------------------------------------------------------------------------------------------
def test_complement_table():
    # Constants.java:11-20,42-49: A<->T, G<->C, '-'->'-', N->N, any other byte < 127 -> 0; reversed
    assert oracle.revcomp(b"AACGTN-").tobytes() == b"-NACGTT"
    assert oracle.revcomp(b"AR*a").tobytes() == b"\x00\x00\x00T"


# --------------------------------------------------------------------------------------------------
# StringIndexer.buildIndexPara — Q1
# --------------------------------------------------------------------------------------------------
def test_index_single_thread():
    o = _o(K=3)
    o.set_reference(b"ACGACGT", index_thread_num=1)
    keys, offs, pos = o.index_csr()
    got = {keys[i].tobytes(): pos[offs[i]:offs[i + 1]].tolist() for i in range(len(keys))}
    assert got == {b"ACG": [0, 3], b"CGA": [1], b"GAC": [2], b"CGT": [4]}


def test_index_q1_duplicate_boundary_position():
    # StringIndexer.java:35,50-58: with 2 threads and |s| = 8, size = 4: task 0 indexes [0,4] INCLUSIVE and
    # task 1 starts at 4, so position 4 is listed twice for its k-mer; lists are concatenated in task order.
    o = _o(K=3)
    o.set_reference(b"ACGTACGT", index_thread_num=2)
    keys, offs, pos = o.index_csr()
    got = {keys[i].tobytes(): pos[offs[i]:offs[i + 1]].tolist() for i in range(len(keys))}
    assert got == {b"ACG": [0, 4, 4], b"CGT": [1, 5], b"GTA": [2], b"TAC": [3]}


# --------------------------------------------------------------------------------------------------
# mergeHits (KMerCounterBase.java:337-450)
# --------------------------------------------------------------------------------------------------
def test_merge_hits_diagonal_chain():
    # read = ref[2:10]: all 6 k-mers (K=3) chain on one diagonal: LongHit(g=2, r=0, len=5); no other 3-mer of
    # the read occurs elsewhere in this reference.
    ref = b"TTACGATCGGAA"
    o = _o(K=3)
    o.set_reference(ref, index_thread_num=1)
    assert o.get_hits(ref[2:10]).tolist() == [[2, 0, 5]]


def test_merge_hits_overlap_truncation_and_drop():
    # ref = ACGTTACGA, read = ACGA (K=3). Raw hits: i=0 "ACG" -> [0,5]; i=1 "CGA" -> [6].
    # Phase 1: LongHit A(0,0,len0), B(5,0) extended by the hit (1,6) to len1. TreeSet order: A, B.
    # Phase 2 (:425-447): A.end = 0+0+3-1 = 2 < 5 -> A emitted whole; B last. => [[0,0,0],[5,0,1]]
    o = _o(K=3)
    o.set_reference(b"ACGTTACGA", index_thread_num=1)
    assert o.get_hits(b"ACGA").tolist() == [[0, 0, 0], [5, 0, 1]]


def test_merge_hits_shorter_overlapping_hit_is_trimmed_away():
    # ref = AAAAAA (K=3), read = AAAA: read k-mers at i=0,1 are both AAA -> positions [0,1,2,3].
    # Phase 1: i=0 creates (0,0),(1,0),(2,0),(3,0); i=1: rHit 0: lHit(0) end+1=1 > 0 -> new (0,1); rHit 1 ==
    # lHit(0).end+1 -> (0,0).len=1; rHit 2 extends (1,0); rHit 3 extends (2,0); (3,0) stays len 0.
    # TreeSet order: (0,0,1) (0,1,0) (1,0,1) (2,0,1) (3,0,0).
    # Phase 2: prev=(0,0,1) end incl = 0+1+3-1 = 3. (0,1,0): overlap, prev.len(1) !< 0 -> lenDiff = 4-0 = 4,
    # curr.len = -4 -> dropped. (1,0,1): not longer -> lenDiff = 3, len = -2 dropped. (2,0,1): lenDiff 2, len -1
    # dropped. (3,0,0): lenDiff 1, len -1 dropped. => [[0,0,1]]
    o = _o(K=3)
    o.set_reference(b"AAAAAA", index_thread_num=1)
    assert o.get_hits(b"AAAA").tolist() == [[0, 0, 1]]


def test_merge_hits_longer_later_hit_truncates_prev():
    # ref = ACGTACGTTT (K=3), read = CGTACGTT.
    # read k-mers: 0 CGT [1,5]; 1 GTA [2]; 2 TAC [3]; 3 ACG [0,4]; 4 CGT [1,5]; 5 GTT [6].
    # chains: (1,0) extended by (1,2),(2,3),(3,4),(4,5),(5,6) -> (1,0,len5); (5,0,len0); (0,3) -> extended by
    # (4,1)? lHit(0,3).end+1 = 1 == 1 -> yes len1, and 5: (5,6)? (0,3,1).end+1=2 != 6 -> stops: (0,3,1).
    # TreeSet: (0,3,1) (1,0,5) (5,0,0).
    # Phase 2: prev=(0,3,1): end incl 0+1+3-1 = 3 >= 1 and prev.len 1 < 5 -> prev.len = 1-0-3 = -2 -> not emitted;
    # prev=(1,0,5): end incl 8 >= 5: (5,0,0) not longer: lenDiff = 9-5 = 4, len -4 dropped. => [[1,0,5]]
    o = _o(K=3)
    o.set_reference(b"ACGTACGTTT", index_thread_num=1)
    assert o.get_hits(b"CGTACGTT").tolist() == [[1, 0, 5]]


def test_merge_hits_q1_duplicate_adds_zero_length_hit():
    # ref ACGTACGT with 2 index threads: "ACG" -> [0,4,4] (see above). read = TACGT (K=3):
    # i=0 TAC [3]; i=1 ACG [0,4,4]; i=2 CGT [1,5].
    # Phase 1: X=(3,0). i=1: rHit 0: X.end+1 = 4 > 0 -> new (0,1); rHit 4 == 4 -> X.len=1, iterator exhausted ->
    # break; the remainder loop adds the second copy as new Z=(4,1). i=2: currSet [(0,1), X(end 4), Z(end 4)]:
    # rHit 1 extends (0,1) -> len1; rHit 5: X.end+1 = 5 -> X.len=2; iterator -> Z; j exhausted -> break.
    # TreeSet: (0,1,1) (3,0,2) (4,1,0). Phase 2: prev (0,1,1) end incl 3 >= 3, prev.len 1 < 2 -> prev.len = 3-0-3 = 0
    # -> emitted (0,1,0); prev = X(3,0,2) end incl 7 >= 4: Z not longer: lenDiff 4, len -4 dropped. => [[0,1,0],[3,0,2]]
    o = _o(K=3)
    o.set_reference(b"ACGTACGT", index_thread_num=2)
    assert o.get_hits(b"TACGT").tolist() == [[0, 1, 0], [3, 0, 2]]
    # with one thread the same read gives the same list here (the extra hit was dropped by the sweep)
    o1 = _o(K=3)
    o1.set_reference(b"ACGTACGT", index_thread_num=1)
    assert o1.get_hits(b"TACGT").tolist() == [[0, 1, 0], [3, 0, 2]]


def test_merge_hits_q2_treeset_key_collision_on_reinsert():
    # Q2 (KMerCounterBase.java:435-441 + Hit.compareTo, Hit.java:13-19): a clipped LongHit is put back with
    # sortedSet.add(curr); when an element with the same (genomePos, readPos) is already in the TreeSet the add is a
    # no-op and the clipped hit is lost, whatever its len.
    # ref (16 bases, all 3-mers distinct) = AACAGATCCGCTGGTA; 2 index threads: size = 8, task 0 indexes [0, 8]
    # inclusive and task 1 starts at 8, so "CGC" -> [8, 8] (Q1).
    # read = ref[2:8] + ref[6:12] = CAGATC TCCGCT. Read k-mers (K = 3): i=0 CAG [2]; 1 AGA [3]; 2 GAT [4]; 3 ATC [5];
    # 4 TCT []; 5 CTC []; 6 TCC [6]; 7 CCG [7]; 8 CGC [8, 8]; 9 GCT [9].
    # Phase 1: P = (2,0) extended at i = 1, 2, 3 -> len 3. i = 4, 5: empty -> the sets are reset. i = 6: X = (6,6).
    # i = 7: 6+0+1 == 7 -> X.len = 1. i = 8: rHit 8 == 6+1+1 -> X.len = 2, the iterator is exhausted -> break; the
    # remainder loop (:412-418) makes the second copy a new hit Z = (8,8,0). i = 9: X: 6+2+1 == 9 -> X.len = 3.
    # TreeSet: P(2,0,3) X(6,6,3) Z(8,8,0).
    # Phase 2: prev = P: 2+3+3-1 = 7 >= 6 and P.len 3 is not < X.len 3 -> lenDiff = 2+3+3-6 = 2, X.len = 1 >= 0,
    # X = (8,8,1), sortedSet.add(X): key (8,8) is Z's -> dropped. Z: 7 >= 8 is false -> emit P, prev = Z; end: emit Z.
    ref = b"AACAGATCCGCTGGTA"
    read = ref[2:8] + ref[6:12]
    o = _o(K=3)
    o.set_reference(ref, index_thread_num=2)
    keys, offs, pos = o.index_csr()
    got = {keys[i].tobytes(): pos[offs[i]:offs[i + 1]].tolist() for i in range(len(keys))}
    assert got[b"CGC"] == [8, 8] and all(len(v) == 1 for k, v in got.items() if k != b"CGC")
    assert o.get_hits(read).tolist() == [[2, 0, 3], [8, 8, 0]]
    # with a single index thread there is no Z: the clipped X = (8,8,1) goes back into the set and survives
    o1 = _o(K=3)
    o1.set_reference(ref, index_thread_num=1)
    assert o1.get_hits(read).tolist() == [[2, 0, 3], [8, 8, 1]]


def test_read_shorter_than_k_and_null_seq():
    o = _o(K=5)
    o.set_reference(b"ACGTACGTACGT", index_thread_num=1)
    assert len(o.get_hits(b"ACG")) == 0 and len(o.seed(b"*")) == 0


# --------------------------------------------------------------------------------------------------
# splitLongHits (KMerCounter.java:142-185)
# --------------------------------------------------------------------------------------------------
def test_split_long_hits_at_contig_boundary():
    # two contigs of 8: ref = ACGTTGCA|GGATCCAT ; the read spans the boundary: read = ref[3:13] (len 10), K=3.
    # One diagonal chain (3,0,len7) spanning 3..12. contigEnd = 8: crosses; first part: 8-3 = 5 >= 3 -> (3,0,len 2);
    # second part: 3+3+7-8 = 5 >= 3 -> readPos += 5, len -= 5, g = 8 -> (8,5,2).
    ref = b"ACGTTGCAGGATCCAT"
    o = _o(K=3)
    o.set_reference(ref, contig_ends=[8, 16], index_thread_num=1)
    h = o.get_hits(ref[3:13]).tolist()
    assert [3, 0, 2] in h and [8, 5, 2] in h
    assert all(not (g < 8 < g + 3 + ln) for g, r, ln in h)  # nothing straddles the boundary any more


# --------------------------------------------------------------------------------------------------
# concordanceArray — Q3, window walk — Q4/Q5, pruneWindows
# --------------------------------------------------------------------------------------------------
def test_window_count_and_prune_by_hand():
    # ref: two copies of the 12-mer W = ACGGTCATTGCA separated by a spacer; read = W; K=4, cutoff 0.
    # Each copy gives one chain (g, 0, len 8). Window of hit i (KMerCounter.java:72-116, 169-191):
    #   start = max(0, g - 0*3/2) = g ; end = min(g + 8 + 4 + (12-0-8-4)*3/2, G) = g + 12 ; count = len = 8.
    # Both exceed the cutoff; they do not overlap -> pruneWindows keeps both.
    W = b"ACGGTCATTGCA"
    ref = W + b"TTTTTTTTTTTTTTTTTTTT" + W
    cut = np.zeros(64, np.int32)
    o = _o(K=4, cutoffs=cut)
    o.set_reference(ref, index_thread_num=1)
    assert o.get_hits(W).tolist() == [[0, 0, 8], [32, 0, 8]]
    assert o.seed(W).tolist() == [[0, 12, 8], [32, 44, 8]]
    # cutoff is strict: count > cutoffs[L]
    cut[12] = 8
    o2 = _o(K=4, cutoffs=cut)
    o2.set_reference(ref, index_thread_num=1)
    assert len(o2.seed(W)) == 0


def test_concordance_bonus_q3_low_coef_zero_vs_msa():
    # read = P + x + Q against ref = P + Q' ... two chains on nearby diagonals. K=4.
    # ref  = ACGGTCAT  TGCAAGCT            (16)
    # read = ACGGTCAT G TGCAAGCT           (17: one inserted base)
    # chains: (0,0,len4) and (8,9,len4). rate = (8-0)/(9-0) = 0.888..: with lowCoef=0 (single-reference Mapper, Q3)
    # 0 <= 0.888 <= 1.25 -> adjacent -> bonus +1; with lowCoef = 1/1.05, highCoef = 1.05 it is NOT adjacent.
    ref = b"ACGGTCATTGCAAGCT"
    read = b"ACGGTCATGTGCAAGCT"
    cut = np.zeros(64, np.int32)
    o = _o(K=4, cutoffs=cut, low_coef=0.0, high_coef=1.25)
    o.set_reference(ref, index_thread_num=1)
    assert o.get_hits(read).tolist() == [[0, 0, 4], [8, 9, 4]]
    # window of hit 0: start 0, end = min(0+4+4+(17-0-4-4)*3/2, 16) = min(21,16) = 16: both hits inside: 4+4+1 = 9
    w = o.window_walk(read)
    assert w[0].tolist() == [0, 16, 9, 9]
    o2 = _o(K=4, cutoffs=cut, low_coef=1 / 1.05, high_coef=1.05)
    o2.set_reference(ref, index_thread_num=1)
    assert o2.window_walk(read)[0].tolist() == [0, 16, 8, 8]


def test_incremental_window_count_equals_direct_recount():
    """The GPU path evaluates every window independently; this is exact because after mergeHits the LongHits
    are sorted and non-overlapping, for which updateCount (KMerCounterBase.java:38-158) equals a recount."""
    from virgena_b200 import synth
    ref, _ = synth.config_reference(1)
    o = _o(K=5, cutoffs=np.zeros(512, np.int32))
    o.set_reference(ref, index_thread_num=8)
    m1, m2 = synth.config_pairs(3, 40)
    n = 0
    for r in list(m1) + [synth.revcomp(x) for x in m2]:
        w = o.window_walk(r)
        assert (w[:, 2] == w[:, 3]).all()
        h = o.get_hits(r)
        assert (np.diff(h[:, 0]) > 0).all() and (h[:-1, 0] + 5 + h[:-1, 2] <= h[1:, 0]).all()
        n += len(w)
    assert n > 10000


def test_init_state_q4_breaks_at_first_non_fitting_hit():
    # Q4 (KMerCounterBase.java:23-27): the loop of getInitState stops at the FIRST hit that does not fit the window,
    # even if a later one would. Crafted list (not something mergeHits can produce: the hits overlap), K = 3, read of 60:
    #   h0 = (10, 0, 2): start = max(0, 10 - 0*3/2) = 10; end = min(10+2+3 + (60-0-2-3)*3/2, 100) = 15 + 165/2 = 97.
    #   h1 = (12, 5, 90): 12+3+90 = 105 > 97 -> break: endIndex = 0, count = len(h0) = 2.
    #   h2 = (20, 10, 1): 20+3+1 = 24 <= 97 would fit, and (20-12)/(10-5) = 1.6 is not adjacent anyway — never looked at.
    w = oracle.walk_crafted([[10, 0, 2], [12, 5, 90], [20, 10, 1]], K=3, read_len=60, from_=0, to=100)
    assert w[0].tolist() == [10, 97, 2, 0, 0]
    # without the long hit in between both fit: count = 2 + 1 (+1: (20-10)/(10-0) = 1.0 lies in [0, 1.25])
    w = oracle.walk_crafted([[10, 0, 2], [20, 10, 1]], K=3, read_len=60, from_=0, to=100)
    assert w[0].tolist() == [10, 97, 4, 0, 1]


def test_update_count_q5_literal_state_machine_on_crafted_list():
    # Q5 (KMerCounterBase.java:38-158): updateCount is an incremental state machine, not a recount. Crafted list
    # (overlapping hits, outside mergeHits' invariant), K = 3, read of 30, window [0, 200), lowCoef 0, highCoef 1.25:
    #   h0 = (50,20,2)  h1 = (52,0,40)  h2 = (58,25,1)
    # isAdjacent[0]: (52-50)/(0-20) = -0.1 -> false; isAdjacent[1]: (58-52)/(25-0) = 0.24 -> true.
    # i=0 getInitState: start = 50 - 20*3/2 = 20; end = 50+2+3 + (30-20-2-3)*3/2 = 55 + 7 = 62; count = 2;
    #     j=1: 52+3+40 = 95 > 62 -> break. (startIndex, endIndex) = (0, 0).
    # i=1 h1: start = 52 - 0 = 52; end = 52+40+3 + (30-0-40-3)*3/2 = 95 + (-39/2 = -19, truncation) = 76.
    #     start moved right (:68-96): h0.genomePos 50 >= 52 is false -> count -= 2 -> 0; the loop over
    #     (startIndex, endIndex] = (0, 0] is empty; j = 1 > endIndex -> scan from 1: h1.genomePos 52 >= 52 -> sIndex = 1.
    #     end moved right (:104-124): prevInWindow = (h[endIndex=0].genomePos 50 >= 52) = false; j = 1: 95 > 76 ->
    #     eIndex = 0. State (52, 76, count 0, startIndex 1, endIndex 0). A recount would see h2 (58 >= 52, 62 <= 76): 1.
    # i=2 h2: start = 58 - 25*3/2 = 58 - 37 = 21; end = 58+1+3 + (30-25-1-3)*3/2 = 62 + 1 = 63.
    #     start moved left (:48-67): startHit = h[1]: prevInWindow = (95 <= 63) = false; j = 0: 50 < 21 no; 55 > 63 no;
    #     no bonus (prevInWindow false); count += 2 -> 2; loop ends, sIndex = 0.
    #     end moved left (:125-152): h[endIndex=0]: 55 > 63 is false -> eIndex stays 0.
    #     State (21, 63, count 2, 0, 0). A recount would give 2 + 1 = 3 (h2 itself lies in its window).
    w = oracle.walk_crafted([[50, 20, 2], [52, 0, 40], [58, 25, 1]], K=3, read_len=30, from_=0, to=200)
    assert w.tolist() == [[20, 62, 2, 0, 0], [52, 76, 0, 1, 0], [21, 63, 2, 0, 0]]


def test_prune_windows_ties_first_wins():
    # pruneWindows (KMerCounterBase.java:317-335): among overlapping windows the larger count wins, ties keep the
    # first in (start asc, end desc) order. Crafted through equal k-mer content at two overlapping loci is hard to
    # do by hand; this is covered by the seeded parity tests. Here: a single window survives a self-overlap.
    W = b"ACGGTCATTGCA"
    o = _o(K=4, cutoffs=np.zeros(64, np.int32))
    o.set_reference(W + W, index_thread_num=1)
    # hits: (0,0,8) and (12,0,8); windows [0,12) and [12,24): touching (end == start) is NOT an overlap (>)
    assert o.seed(W).tolist() == [[0, 12, 8], [12, 24, 8]]


# --------------------------------------------------------------------------------------------------
# SmithWatermanGotoh (SmithWatermanGotoh.java:51-222)
# --------------------------------------------------------------------------------------------------
def test_sw_identical():
    rec, seq = _o().sw_align(b"ACGTACGT", b"ACGTACGT")
    assert rec.tolist() == [16, 0, 8, 0, 8, 8, 8] and seq == b"ACGTACGT"


def test_sw_q8_no_positive_cell():
    # no cell > 0: Cell stays (0,0), score 0, empty alignment (:146-150, :178-221)
    rec, seq = _o().sw_align(b"AAAA", b"CCCC")
    assert rec.tolist() == [0, 0, 0, 0, 0, 0, 0] and seq == b""


def test_sw_mismatch_inside():
    # ACGTTACGT vs ACGTAACGT with 2/-3/5/2: through the mismatch 4*2 - 3 + 4*2 = 13 > 8 (either flank alone);
    # a gap would cost 5. identity = 8, sequence1 carries the read base at the mismatch.
    rec, seq = _o().sw_align(b"ACGTTACGT", b"ACGTAACGT")
    assert rec.tolist() == [13, 0, 9, 0, 9, 8, 9] and seq == b"ACGTTACGT"


def test_sw_deletion_in_read_is_gap_run_with_affine_cost():
    # read lacks 2 bases of the window: 10 matches on each side: 20*2 - (5 + 2) = 33; '-' x 2 in sequence1.
    a, b = b"ACGGTCATTG", b"CAAGCTGGAT"
    rec, seq = _o().sw_align(a + b, a + b"TT" + b)
    assert rec[0] == 33 and seq == a + b"--" + b
    assert rec.tolist() == [33, 0, 20, 0, 22, 20, 22]


def test_sw_insertion_in_read_is_lower_case():
    a, b = b"ACGGTCATTG", b"CAAGCTGGAT"
    rec, seq = _o().sw_align(a + b"TTT" + b, a + b)
    # 20 matches - (5 + 2*2) = 31; inserted read bases are lower-cased (Constants.lower, :191)
    assert rec.tolist() == [31, 0, 23, 0, 20, 20, 23] and seq == a + b"ttt" + b


def test_sw_first_maximum_in_row_major_order():
    # two equally good local alignments: the first strictly-greater cell in row-major order wins (:146-150).
    # s1 = ACGT, s2 = ACGTNNACGT: row 4 reaches 8 first at column 4 -> end2 = 4.
    rec, seq = _o().sw_align(b"ACGT", b"ACGTNNACGT")
    assert rec.tolist() == [8, 0, 4, 0, 4, 4, 4]
    # s1 = ACGTNNACGT, s2 = ACGT: 8 is reached in row 4 (col 4) before row 10
    rec, seq = _o().sw_align(b"ACGTTTACGT", b"ACGT")
    assert rec.tolist() == [8, 0, 4, 0, 4, 4, 4]


def test_sw_pointer_priority_on_ties():
    # SmithWatermanGotoh.java:135-143: STOP (v == 0) > DIAGONAL (v == f) > UP (v == g) > LEFT. Three cases where the
    # tie decides the alignment; V matrices worked out by hand (rows = s1, columns = s2).
    # (a) v == f == g with 2/-1/3/1, s1 = AAGCA, s2 = AACAGGA:
    #        A  A  C  A  G  G  A
    #     A  2  2  0  2  0  0  2
    #     A  2  4  1  2  1  0  2
    #     G  0  1  3  0  4  3  0
    #     C  0  0  3  2  1  3  2
    #     A  2  2  0  5  2  1  5      first maximum 5 at (5,4)
    #     path (5,4) D (4,3) D (3,2): f = V(2,1) + mismatch = 2 - 1 = 1 and g = V(2,2) - gop = 4 - 3 = 1: DIAGONAL wins
    #     -> (2,1) D -> (1,0) STOP: sequence1 = AGCA, start1 = 1, start2 = 0, identity 3 (G/A mismatches).
    #     UP first would give AAgCA.
    rec, seq = _o(match=2, mismatch=-1, gop=3, gep=1).sw_align(b"AAGCA", b"AACAGGA")
    assert rec.tolist() == [5, 1, 5, 0, 4, 3, 4] and seq == b"AGCA"
    # (b) v == f == h with the default 2/-3/5/2, s1 = CCCAGCAA, s2 = CCCGAGCA: maximum 9 at (7,8); the path runs
    #     diagonally to (4,5) = 3, then (3,4): f = V(2,3) + mismatch(C/G) = 4 - 3 = 1 and h = V(3,3) - gop = 6 - 5 = 1:
    #     DIAGONAL wins -> (2,3)=4, (1,2)=2, (0,1) STOP: CCCAGCA, start1 0, start2 1; LEFT first would give CCC-AGCA.
    rec, seq = _o().sw_align(b"CCCAGCAA", b"CCCGAGCA")
    assert rec.tolist() == [9, 0, 7, 1, 8, 6, 7] and seq == b"CCCAGCA"
    # (c) v == g == h with 1/-1/1/1, s1 = CACAGA, s2 = ACAAGAA: maximum 4 at (6,6), diagonal down to (3,3) = 1 where
    #     f = V(2,2) + mismatch = 0 - 1, g = V(2,3) - gop = 2 - 1 = 1, h = V(3,2) - gop = 2 - 1 = 1: UP wins over LEFT
    #     -> one lower-case read base, then (2,3) D (1,2) D (0,1): CAcAGA; LEFT first would give AC-AGA.
    rec, seq = _o(match=1, mismatch=-1, gop=1, gep=1).sw_align(b"CACAGA", b"ACAAGAA")
    assert rec.tolist() == [4, 0, 6, 1, 6, 5, 6] and seq == b"CAcAGA"


def test_sw_q9_unknown_base_in_insertion_becomes_zero_byte():
    a, b = b"ACGGTCATTG", b"CAAGCTGGAT"
    rec, seq = _o().sw_align(a + b"R" + b, a + b)
    assert seq == a + b"\x00" + b  # Constants.lower has no entry for 'R' (:22-30)


def test_sw_n_matches_n():
    rec, seq = _o().sw_align(b"ACNGT", b"ACNGT")
    assert rec[0] == 10 and rec[5] == 5  # byte equality: N == N scores as a match (:108)


def test_find_max_score_equals_align_score():
    o = _o()
    for a, b in ((b"ACGTTACGT", b"ACGTAACGT"), (b"AAAA", b"CCCC"), (b"GATTACA", b"GCATGCT")):
        assert o.sw_max_score(a, b) == o.sw_align(a, b)[0][0]


# --------------------------------------------------------------------------------------------------
# pileup / consensus / CIGAR
# --------------------------------------------------------------------------------------------------
def _rec(start, start2, seq1):
    r = np.zeros(1, oracle.PAIR_DTYPE)
    r["m"]["seq1_offset"] = -1
    m = r["m"][0, 0]
    m["present"], m["start"], m["start2"], m["length"], m["seq1_offset"] = 1, start, start2, len(seq1), 0
    r["m"][0, 0] = m
    return r


def test_pileup_q11_and_insertions():
    # ConsensusBuilderSimple.java:56-75: lower case skipped, '-' -> column 4, N (and any unlisted byte) -> A.
    counts = oracle.pileup(_rec(2, 1, b"ANt-G"), np.frombuffer(b"ANt-G", np.uint8), 10)
    exp = np.zeros((10, 5), np.int32)
    exp[3, 0] = 1   # A at 2+1
    exp[4, 0] = 1   # N -> A (Q11)
    exp[5, 4] = 1   # '-' ; the insertion 't' does not advance
    exp[6, 2] = 1   # G
    assert (counts == exp).all()


def test_consensus_argmax_first_max_wins_and_threshold():
    counts = np.zeros((4, 5), np.int32)
    counts[0] = [2, 2, 0, 0, 9]   # tie A/T -> A (strict >); the gap column is ignored
    counts[1] = [0, 1, 3, 3, 0]   # tie G/C -> G
    counts[2] = [0, 0, 0, 1, 0]   # 1 <= threshold 1 -> uncovered
    counts[3] = [0, 0, 0, 5, 0]
    assert oracle.consensus(counts, b"TTTT", False, 1).tobytes() == b"AG-C"
    assert oracle.consensus(counts, b"TTTT", True, 1).tobytes() == b"AGTC"


def test_cigar_and_xn():
    # BamPrinter.setCigar (BamPrinter.java:121-171)
    assert oracle.cigar(b"ACGT--ACgtAC", 2, 10, 12, 7) == ("2S4M2D2M2I2M2S", 1)
    # an alignment cannot start with an insertion or deletion in practice; a leading I would emit "0M" first
    assert oracle.cigar(b"aCG", 0, 3, 3, 2)[0] == "0M1I2M"


def test_map_read_to_region_hand_derived():
    # KMerCounter.mapReadToRegion (KMerCounter.java:119-140) / computeKMerCount (KMerCounterBase.java:468-487).
    # ref = AAAAACCCCCGGGGGTTTTT, K = 5, read = CCCCCGGGGG: read k-mer i occurs only at reference position 5 + i, i.e. one
    # chain LongHit(g=5, r=0, len=5); window start = max(0, 5 - 0*3/2) = 5, end = min(5 + 5 + 5 + (10-0-5-5)*3/2, 20) = 15,
    # count = len = 5.
    o = oracle.Oracle(K=5, cutoffs=np.zeros(64, np.int32))
    o.set_reference(b"AAAAACCCCCGGGGGTTTTT", index_thread_num=1)
    assert o.map_read_to_region(b"CCCCCGGGGG") == (5, 15, 5)
    # CCCCCTTTTT: two isolated zero-length hits (5, 0) and (15, 5). Hit 0: start 5, end = min(5 + 0 + 5 + (10-0-0-5)*3/2, 20)
    # = 17 (integer 15/2 = 7), hit 1 ends at 20 > 17: count 0. Hit 1: window [8, 20) holds only itself: count 0, not
    # strictly greater: the first window stays.
    assert o.map_read_to_region(b"CCCCCTTTTT") == (5, 17, 0)
    assert o.map_read_to_region(b"ACGTACGTAC") is None         # no k-mer of the read occurs in the reference


def test_bam_record_fields():
    # BamPrinter.setRecordFieldsMapped / setRecordFieldsUnmapped (BamPrinter.java:16-115), hand-derived:
    # contigs [0,100) [100,250); pair 0 concordant (mate1 forward at 95+8=103 -> contig 1, pos 4; mate2 reverse at 20+0),
    # pair 1 left-mate-mapped, pair 2 unmapped.
    import virgena_b200 as vg
    rec = np.zeros(3, vg.PAIR_DTYPE)
    rec["m"]["seq1_offset"] = -1
    pool = np.frombuffer(b"ACGTaC-G" + b"TTTT", np.uint8)
    rec["cls"] = [0, 3, 5]
    m = rec["m"][0][0]
    m["present"], m["start"], m["start2"], m["reverse"], m["start1"], m["end1"], m["identity"], m["length"], m["seq1_offset"] = 1, 95, 8, 0, 1, 8, 5, 8, 0
    m = rec["m"][0][1]
    m["present"], m["start"], m["start2"], m["reverse"], m["start1"], m["end1"], m["identity"], m["length"], m["seq1_offset"] = 1, 20, 0, 1, 0, 4, 4, 4, 8
    m = rec["m"][1][0]
    m["present"], m["start"], m["start2"], m["reverse"], m["start1"], m["end1"], m["identity"], m["length"], m["seq1_offset"] = 1, 240, 20, 0, 0, 4, 3, 4, 8
    out = oracle.bam_records(rec, pool, [10, 4, 5], [4, 6, 5], [100, 250])
    a, b, c, d, e, f = out
    assert (a["flag"], a["ref_index"], a["pos"], a["mapq"], a["mate_ref_index"], a["mate_pos"], a["xn"]) == (0x1 | 0x2 | 0x20 | 0x40, 1, 4, 255, 0, 21, 1)
    assert vg.cigar_string(a["cigar"]) == "1S4M1I1M1D1M2S"
    assert (b["flag"], b["ref_index"], b["pos"], b["mate_ref_index"], b["mate_pos"], b["xn"]) == (0x1 | 0x2 | 0x10 | 0x80, 0, 21, 1, 4, 0)
    assert vg.cigar_string(b["cigar"]) == "4M"
    # position 260 lies behind the last contig: the Java loop runs off the end (refIndex = 2, no offset subtracted)
    assert (c["flag"], c["ref_index"], c["pos"], c["mate_ref_index"], c["mate_pos"], c["xn"]) == (0x1 | 0x8 | 0x40, 2, 261, -1, 0, 1)
    assert (d["flag"], d["ref_index"], d["pos"], d["mapq"], d["mate_ref_index"], d["mate_pos"]) == (0x1 | 0x4 | 0x80, -1, 0, 0, 2, 261)
    assert vg.cigar_string(d["cigar"]) == "*"
    assert (e["flag"], f["flag"]) == (0x1 | 0x4 | 0x8 | 0x40, 0x1 | 0x4 | 0x8 | 0x80)


# --------------------------------------------------------------------------------------------------
# Mapper.mapRead pairing (Mapper.java:92-390)
# --------------------------------------------------------------------------------------------------
def test_map_read_classes():
    from virgena_b200 import synth
    ref, _ = synth.config_reference(1)
    cut = np.full(400, 100, np.int32)  # above any random window (~70), below a true one (~240)
    o = _o(K=5, cutoffs=cut)
    o.set_reference(ref, index_thread_num=1)
    f = ref[1000:1250].tobytes()
    r = synth.revcomp(ref[1300:1550]).tobytes()
    junk = bytes(np.random.default_rng(0).integers(65, 70, 250).astype(np.uint8))
    far = synth.revcomp(ref[6000:6250]).tobytes()
    r1 = [f, r, f, junk, junk, f, b"*"]
    r2 = [r, f, junk, r, junk, far, r]
    b = synth.pack_ragged(r1, r2)
    rec, pool, st = o.map_pairs(*b)
    # FR concordant, RF concordant, left only, right only, unmapped, discordant (insert > 1000), right only
    assert rec["cls"].tolist() == [0, 1, 3, 4, 5, 2, 4]
    assert rec["m"]["reverse"][0].tolist() == [0, 1] and rec["m"]["reverse"][1].tolist() == [1, 0]
    # perfect 250-mers: score 500, identity 250, full-length alignment inside the candidate window
    m = rec["m"][0, 0]
    assert m["score"] == 500 and m["identity"] == 250 and m["start"] + m["start2"] == 1000
    # stats 1)-7): pairs, bothExist, readsTotal, totalMapped, fwd, rev, bothMapped, totalScore, concordant
    assert st[0] == 7 and st[1] == 6 and st[2] == 13 and st[3] == 9 and st[6] == 3 and st[8] == 2


def test_pair_select_q7_fragment_id_stays_zero():
    # Q7 (MappedRead.chooseFragment, MappedRead.java:39-50): a window that fits no fragment keeps fragmentID = 0 (the
    # field's default), so it pairs with mates in fragment 0. Fragments [0,100) [100,200) [200,300), insert 1000.
    # reverse1 = [(190, 210, 30)] straddles fragments 1 / 2 -> fragmentID 0; forward2 = [(10, 60, 20)] lies in fragment 0.
    # Side 1 (Mapper.java:209-224): read2.start 10 < read1.end 210, 210 - 10 = 200 <= 1000, fragment 0 == 0 ->
    # concordant RF across two fragments.
    fe = [100, 200, 300]
    d = oracle.pair_select([], [[190, 210, 30]], [[10, 60, 20]], [], fe, True)
    assert d == dict(cls=1, r1=(1, 0), r2=(0, 0), frag1=0, frag2=0)
    # the same mate properly inside fragment 1: fragment 1 != 0 -> no concordant pair; best singles exist on both sides, and
    # the fragment-aware rule (:309-353) finds no fragment holding both mates -> the pair is UNMAPPED (:347-351)
    d = oracle.pair_select([], [[150, 190, 30]], [[10, 60, 20]], [], fe, True)
    assert d == dict(cls=5, r1=None, r2=None, frag1=-1, frag2=-1)
    # not fragmented: plain discordant with the best singles (:355-385)
    d = oracle.pair_select([], [[150, 190, 30]], [[10, 60, 20]], [], fe, False)
    assert d["cls"] == 1  # 10 < 190 and 190 - 10 <= 1000 with fragmentID 0 == 0 (chooseFragment is never called)


def test_pair_select_fragment_aware_discordant_rule():
    # Mapper.java:309-353: when no concordant pair exists in a fragmented genome, the pair goes to the fragment with the
    # largest countFrag1 + countFrag2 among those where BOTH mates have a candidate — not to the best singles.
    # forward1 = [(10,60,50) frag 0, (110,160,30) frag 1]; forward2 = [(120,170,25) frag 1]; reverse2 = [(210,260,45) frag 2].
    # Side 0: forward1 x reverse2: fragments 0 != 2 and 1 != 2 -> nothing; side 1: reverse1 is empty.
    # Best singles: mate 1 (10,60,50); mate 2 scans reverse first (:158-191): 45, forward 25 is not greater.
    # countFrag1 = [50,30,0], countFrag2 = [0,25,45] -> only fragment 1 has both: r1 = forward1[1], r2 = forward2[0].
    fe = [100, 200, 300]
    f1, f2, r2 = [[10, 60, 50], [110, 160, 30]], [[120, 170, 25]], [[210, 260, 45]]
    d = oracle.pair_select(f1, [], f2, r2, fe, True)
    assert d == dict(cls=2, r1=(0, 1), r2=(0, 0), frag1=1, frag2=1)
    # the unfragmented rule would take the best singles: forward1[0] and reverse2[0] — but here they even satisfy side 0
    # (10 < 260, 250 <= 1000, fragmentID 0 == 0 without chooseFragment) and come out concordant
    assert oracle.pair_select(f1, [], f2, r2, fe, False) == dict(cls=0, r1=(0, 0), r2=(1, 0), frag1=0, frag2=0)
    # strict '>' everywhere: equal sums keep the first fragment, equal counts keep the first candidate, and mate 2
    # prefers its REVERSE list on ties because that list is scanned first
    d = oracle.pair_select([[10, 60, 20], [110, 160, 20]], [], [[20, 70, 20]], [[120, 170, 20]], fe, True, insert_len=5)
    # side 0: (10,60) x (120,170): fragments differ; (110,160) x (120,170): 110 < 170 but 170 - 110 = 60 > 5 -> none.
    # countFrag1 = [20,20,0], countFrag2 = [20,20,0]: fragment 0 reaches 40 first; fragment 1 is not greater.
    assert d == dict(cls=2, r1=(0, 0), r2=(0, 0), frag1=0, frag2=0)


def test_identity_counting_hand_derived():
    """ProblemIntervalBuilder.IdentityCounting.run (ProblemIntervalBuilder.java:35-133), three walks derived by hand:
    a mismatch leaving and entering the window; a soft-clipped read head (the clipped bases count in len, not in identity);
    an insertion (lower case: len grows / shrinks as it enters / leaves) and a deletion ('-' never matches)."""
    import oracle

    def one(genome, seq1, mate, read_len, window, **f):
        rec = np.zeros(1, oracle.PAIR_DTYPE)
        m = rec["m"][0][mate]
        m["present"], m["start"], m["seq1_offset"], m["length"] = 1, 0, 0, len(seq1)
        for k, v in f.items():
            m[k] = v
        lens = [[read_len], [0]] if mate == 0 else [[0], [read_len]]
        return oracle.identity_counting(rec, np.frombuffer(seq1, np.uint8), lens[0], lens[1], np.frombuffer(genome, np.uint8), window, 1)

    ident, cov = one(b"ACGTACGTAC", b"GAACGT", 0, 6, 4, start1=0, end1=6, start2=2, end2=8, identity=5)
    assert ident.tolist() == [0, 0, 0.75, 0.75, 1.0, 0, 0, 0, 0, 0] and cov.tolist() == [0, 0, 1, 1, 1, 0, 0, 0, 0, 0]
    ident, cov = one(b"ACGTACGTAC", b"GTACGT", 1, 8, 4, start1=2, end1=8, start2=2, end2=8, identity=6)
    assert ident.tolist() == [0.5, 0.75, 1.0, 1.0, 1.0, 0, 0, 0, 0, 0] and cov.tolist() == [1, 1, 1, 1, 1, 0, 0, 0, 0, 0]
    ident, cov = one(b"ACGTACGTACGT", b"GTaA-GT", 0, 6, 3, start1=0, end1=6, start2=2, end2=8, identity=5)
    third = np.float32(2) / np.float32(3)
    assert ident.tolist() == [0, 0, 0.75, 0.5, third, third, 0, 0, 0, 0, 0, 0] and cov.tolist() == [0, 0, 1, 1, 1, 1] + [0] * 6
    # reads shorter than the window are skipped but still count when the list is cut into per-thread chunks
    ident, cov = one(b"ACGTACGTAC", b"GAACGT", 0, 6, 7, start1=0, end1=6, start2=2, end2=8, identity=5)
    assert not ident.any() and not cov.any()
