"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/oracle.cpp header).

ctypes binding of liboracle.so, the C++ restatement of VirGenA's Java hot path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module. The product (virgena_b200) never does.

PARITY UNPINNED BY THE REFERENCE: no golden vectors exist upstream and the
reference cannot run here; the oracle is pinned by hand-derived KATs
(tests/test_oracle_kat.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MATE_DTYPE = np.dtype(
    [
        ("present", "<i4"), ("start", "<i4"), ("end", "<i4"), ("count", "<i4"),
        ("reverse", "<i4"), ("fragment_id", "<i4"),
        ("score", "<i4"), ("start1", "<i4"), ("end1", "<i4"), ("start2", "<i4"), ("end2", "<i4"),
        ("identity", "<i4"), ("length", "<i4"), ("pad_", "<i4"),
        ("seq1_offset", "<i8"),
    ],
    align=True,
)
PAIR_DTYPE = np.dtype([("cls", "<i4"), ("pad_", "<i4"), ("m", MATE_DTYPE, (2,))], align=True)
assert MATE_DTYPE.itemsize == 64 and PAIR_DTYPE.itemsize == 136


def build(force=False):
    """Compile liboracle.so with the committed Makefile (g++ only)."""
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.vo_create.restype = C.c_void_p
        L.vo_create.argtypes = [C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.vo_destroy.argtypes = [C.c_void_p]
        L.vo_set_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.vo_index_num_keys.argtypes = [C.c_void_p, C.c_int]
        L.vo_index_num_positions.restype = C.c_int64
        L.vo_index_num_positions.argtypes = [C.c_void_p, C.c_int]
        L.vo_index_export.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.vo_set_msa.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_int]
        L.vo_get_hits.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.vo_seed.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.vo_window_walk.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.vo_sw_align.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.vo_sw_max_score.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.vo_sw_batch.argtypes = [C.c_int] * 4 + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.vo_map_pairs.restype = C.c_int64
        L.vo_map_pairs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.vo_pileup.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
        L.vo_consensus.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.vo_cigar.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.vo_revcomp.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.vo_cutoffs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_int,
                                 C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(x):
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8)
    if isinstance(x, str):
        return np.frombuffer(x.encode("latin1"), dtype=np.uint8)
    return np.ascontiguousarray(x, dtype=np.uint8)


def _i32(x):
    return np.ascontiguousarray(x, dtype=np.int32)


class Oracle:
    """Mapper(document) + Reference of the Java, on the CPU."""

    def __init__(self, K=5, low_coef=0.0, high_coef=1.25, cutoffs=None, insert_len=1000,
                 match=2, mismatch=-3, gop=5, gep=2):
        self.K = K
        self.cutoffs = _i32(cutoffs if cutoffs is not None else np.zeros(1024, np.int32))
        self.sw = (match, mismatch, gop, gep)
        self.h = lib().vo_create(K, low_coef, high_coef, _p(self.cutoffs), len(self.cutoffs), insert_len,
                                 match, mismatch, gop, gep)
        self.G = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().vo_destroy(self.h)
            self.h = None

    def set_reference(self, seq, contig_ends=None, fragment_ends=None, is_fragmented=False, index_thread_num=1):
        seq = _u8(seq)
        self.G = len(seq)
        ce = _i32(contig_ends if contig_ends is not None else [len(seq)])
        fe = _i32(fragment_ends if fragment_ends is not None else ce)
        lib().vo_set_reference(self.h, _p(seq), len(seq), _p(ce), len(ce), _p(fe), len(fe), int(is_fragmented),
                               index_thread_num)

    def set_msa(self, rows, K=7, low_coef=1 / 1.5, high_coef=1.5, cutoffs=None):
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        cut = _i32(cutoffs if cutoffs is not None else np.zeros(1024, np.int32))
        self.K_msa = K
        lib().vo_set_msa(self.h, _p(rows), rows.shape[0], rows.shape[1], K, low_coef, high_coef, _p(cut), len(cut))

    def index_csr(self, msa=False):
        K = self.K_msa if msa else self.K
        nk = lib().vo_index_num_keys(self.h, int(msa))
        npos = lib().vo_index_num_positions(self.h, int(msa))
        keys = np.zeros((nk, K), np.uint8)
        offs = np.zeros(nk + 1, np.int64)
        pos = np.zeros(max(npos, 1), np.int32)
        lib().vo_index_export(self.h, int(msa), _p(keys), _p(offs), _p(pos))
        return keys, offs, pos[:npos]

    def get_hits(self, read, msa=False, split=True):
        read = _u8(read)
        out = np.zeros((8192, 3), np.int32)
        n = lib().vo_get_hits(self.h, int(msa), _p(read), len(read), _p(out), len(out), int(split and not msa))
        assert n <= len(out)
        return out[:n].copy()

    def map_read_to_region(self, read, msa=False):
        """KMerCounter.mapReadToRegion(read, index, 0, length): (start, end, count) or None; count = computeKMerCount."""
        read = _u8(read)
        out = np.zeros(3, np.int32)
        lib().vo_map_read_to_region.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        n = lib().vo_map_read_to_region(self.h, int(msa), _p(read if len(read) else np.zeros(1, np.uint8)), len(read), _p(out))
        assert n >= 0, "mapReadToRegion and computeKMerCount disagree"
        return tuple(int(x) for x in out) if n else None

    def seed(self, read, msa=False):
        read = _u8(read)
        out = np.zeros((4096, 3), np.int32)
        n = lib().vo_seed(self.h, int(msa), _p(read), len(read), _p(out), len(out))
        assert n <= len(out)
        return out[:n].copy()

    def window_walk(self, read, msa=False):
        read = _u8(read)
        out = np.zeros((8192, 4), np.int32)
        n = lib().vo_window_walk(self.h, int(msa), _p(read), len(read), _p(out), len(out))
        return out[:n].copy()

    def sw_align(self, s1, s2):
        s1, s2 = _u8(s1), _u8(s2)
        rec = np.zeros(7, np.int32)
        buf = np.zeros(len(s1) + len(s2) + 1, np.uint8)
        n = lib().vo_sw_align(*self.sw, _p(s1), len(s1), _p(s2), len(s2), _p(rec), _p(buf), len(buf))
        return rec, bytes(buf[:n])

    def sw_max_score(self, s1, s2):
        s1, s2 = _u8(s1), _u8(s2)
        return lib().vo_sw_max_score(*self.sw, _p(s1), len(s1), _p(s2), len(s2))

    def sw_batch(self, s1_pool, s1_off, s2_pool, s2_off, n_threads=1, want_seq=True):
        n = len(s1_off) - 1
        s1_off = np.ascontiguousarray(s1_off, np.int64)
        s2_off = np.ascontiguousarray(s2_off, np.int64)
        recs = np.zeros((n, 7), np.int32)
        cap = (np.diff(s1_off) + np.diff(s2_off)).astype(np.int64)
        soff = np.zeros(n + 1, np.int64)
        np.cumsum(cap, out=soff[1:])
        pool = np.zeros(int(soff[-1]) + 1, np.uint8) if want_seq else None
        lib().vo_sw_batch(*self.sw, _p(_u8(s1_pool)), _p(s1_off), _p(_u8(s2_pool)), _p(s2_off), n, _p(recs),
                          _p(pool), _p(soff), n_threads)
        return recs, pool, soff

    def map_pairs(self, bases, off1, len1, off2, len2, msa=False, batch_size=1000, n_threads=1):
        bases = _u8(bases)
        off1 = np.ascontiguousarray(off1, np.int64)
        off2 = np.ascontiguousarray(off2, np.int64)
        len1, len2 = _i32(len1), _i32(len2)
        n = len(off1)
        out = np.zeros(n, PAIR_DTYPE)
        cap = int((len1.astype(np.int64).sum() + len2.astype(np.int64).sum()) * 3) + 16
        pool = np.zeros(cap, np.uint8)
        stats = np.zeros(9, np.int64)
        used = lib().vo_map_pairs(self.h, int(msa), _p(bases), _p(off1), _p(off2), _p(len1), _p(len2), n, batch_size,
                                  n_threads, _p(out), _p(pool), cap, _p(stats))
        assert used <= cap
        return out, pool[:used].copy(), stats


def pileup(recs, pool, G, counts=None):
    counts = np.zeros((G, 5), np.int32) if counts is None else counts
    rc = lib().vo_pileup(_p(recs), len(recs), _p(_u8(pool) if len(pool) else np.zeros(1, np.uint8)), G, _p(counts))
    if rc != 0:
        raise ValueError("pileup: byte in ('T','a') or position past the genome (Java would throw)")
    return counts


def identity_counting(recs, pool, len1, len2, genome, window_size, thread_num):
    """ProblemIntervalBuilder.countIdentity: (identity float32[G], coverage int32[G]) over the mapped reads of `recs`."""
    genome = _u8(genome)
    G = len(genome)
    ident = np.zeros(max(G, 1), np.float32)
    cov = np.zeros(max(G, 1), np.int32)
    L = lib()
    L.vo_identity_counting.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p]
    L.vo_identity_counting.restype = C.c_int
    rc = L.vo_identity_counting(_p(recs), len(recs), _p(_u8(pool) if len(pool) else np.zeros(1, np.uint8)), _p(_i32(len1)),
                                _p(_i32(len2)), _p(genome if G else np.zeros(1, np.uint8)), G, window_size, thread_num,
                                _p(ident), _p(cov))
    if rc != 0:
        raise ValueError("identity counting: the Java would throw ArrayIndexOutOfBounds")
    return ident[:G], cov[:G]


def consensus(counts, genome, save_uncovered=False, coverage_threshold=0):
    genome = _u8(genome)
    counts = np.ascontiguousarray(counts, np.int32)
    out = np.zeros(len(genome), np.uint8)
    lib().vo_consensus(_p(counts), len(genome), int(save_uncovered), coverage_threshold, _p(genome), _p(out))
    return out


def cigar(seq1, start1, end1, read_len, identity):
    seq1 = _u8(seq1)
    ops = np.zeros(len(seq1) + 4, np.uint32)
    xn = C.c_int(0)
    n = lib().vo_cigar(_p(seq1), len(seq1), start1, end1, read_len, identity, _p(ops), len(ops), C.byref(xn))
    s = "".join("%d%s" % (int(o) >> 4, "MID.S"[int(o) & 15]) for o in ops[:n])
    return s, xn.value


def cigar_ops(seq1, start1, end1, read_len, identity):
    """setCigar (BamPrinter.java:121-171) as (array of len << 4 | op, XN)."""
    seq1 = _u8(seq1)
    ops = np.zeros(len(seq1) + 4, np.uint32)
    xn = C.c_int(0)
    n = lib().vo_cigar(_p(seq1 if len(seq1) else np.zeros(1, np.uint8)), len(seq1), start1, end1, read_len, identity, _p(ops),
                       len(ops), C.byref(xn))
    return ops[:n].copy(), xn.value


def bam_records(recs, pool, len1, len2, contig_ends):
    """BamPrinter.setRecordFieldsMapped / setRecordFieldsUnmapped (BamPrinter.java:16-115) + printBAM's proper-pair flag
    (:259-317) for every pair result, two records per pair in (pair, mate) order. recs = the oracle's / library's pair
    records, pool = sequence1 bytes, len1/len2 = read lengths per pair. Returns a list of dicts with the same fields as
    vg_bam_record and the CIGAR elements as an array (empty = "*")."""
    contig_ends = [int(x) for x in contig_ends]

    def locate(s):  # BamPrinter.java:21-30 (runs off the end when s lies behind the last contig)
        prev = 0
        r = 0
        while r < len(contig_ends):
            if s < contig_ends[r]:
                s -= prev
                break
            prev = contig_ends[r]
            r += 1
        return r, s + 1

    out = []
    for pi in range(len(recs)):
        cls = int(recs["cls"][pi])
        proper = cls in (0, 1)  # VG_CLS_CONCORDANT_FR / _RF
        for mate in range(2):
            me, other = recs["m"][pi][mate], recs["m"][pi][1 - mate]
            d = dict(flag=0x1 | (0x80 if mate else 0x40) | (0x2 if proper else 0), ref_index=-1, pos=0, mapq=0, mate_ref_index=-1,
                     mate_pos=0, xn=0, cigar=np.zeros(0, np.uint32))
            if me["present"]:
                d["ref_index"], d["pos"] = locate(int(me["start"]) + int(me["start2"]))
                d["mapq"] = 255
                if me["reverse"]:
                    d["flag"] |= 0x10
                off, ln = int(me["seq1_offset"]), int(me["length"])
                rl = int(len2[pi] if mate else len1[pi])
                d["cigar"], d["xn"] = cigar_ops(pool[off:off + ln], int(me["start1"]), int(me["end1"]), rl, int(me["identity"]))
            else:
                d["flag"] |= 0x4
            if other["present"]:
                d["mate_ref_index"], d["mate_pos"] = locate(int(other["start"]) + int(other["start2"]))
                if other["reverse"]:
                    d["flag"] |= 0x20
            else:
                d["flag"] |= 0x8
            out.append(d)
    return out


def revcomp(seq):
    seq = _u8(seq)
    out = np.zeros(len(seq), np.uint8)
    lib().vo_revcomp(_p(seq), len(seq), _p(out))
    return out


def cutoffs(seq, K=5, coef=1.25, order=4, read_num=10000, min_len=250, max_len=250, step=10, p_value=0.01,
            contig_ends=None, is_fragmented=False, index_thread_num=1, seed=1, msa_rows=None):
    seq = _u8(seq)
    ce = _i32(contig_ends if contig_ends is not None else [len(seq)])
    out = np.zeros(max_len + 1, np.int32)
    if msa_rows is not None:
        rows = np.ascontiguousarray(msa_rows, np.uint8)
        lib().vo_cutoffs(_p(seq), len(seq), _p(ce), len(ce), int(is_fragmented), K, coef, order, read_num, min_len,
                         max_len, step, p_value, index_thread_num, seed, 1, _p(rows), rows.shape[0], rows.shape[1], _p(out))
    else:
        lib().vo_cutoffs(_p(seq), len(seq), _p(ce), len(ce), int(is_fragmented), K, coef, order, read_num, min_len,
                         max_len, step, p_value, index_thread_num, seed, 0, None, 0, 0, _p(out))
    return out


def random_model(seq, K=5, coef=1.25, order=4, read_num=10000, min_len=250, max_len=250, step=10, p_value=0.01,
                 contig_ends=None, is_fragmented=False, index_thread_num=1, seed=1, msa_rows=None):
    """RandomModel.genModel + computeCuttofs with the seeded stand-in RNG: (cutoffs, sorted counts [size][read_num],
    the read_num reads of length max_len)."""
    seq = _u8(seq)
    ce = _i32(contig_ends if contig_ends is not None else [len(seq)])
    out = np.zeros(max_len + 1, np.int32)
    size = (max_len - min_len) // step + 1
    counts = np.zeros((size, read_num), np.int32)
    reads = np.zeros((read_num, max_len), np.uint8)
    f = lib().vo_random_model
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                  C.c_int, C.c_float, C.c_int, C.c_uint64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                  C.c_void_p]
    if msa_rows is not None:
        rows = np.ascontiguousarray(msa_rows, np.uint8)
        f(_p(seq), len(seq), _p(ce), len(ce), int(is_fragmented), K, coef, order, read_num, min_len, max_len, step, p_value,
          index_thread_num, seed, 1, _p(rows), rows.shape[0], rows.shape[1], _p(out), _p(counts), _p(reads))
    else:
        f(_p(seq), len(seq), _p(ce), len(ce), int(is_fragmented), K, coef, order, read_num, min_len, max_len, step, p_value,
          index_thread_num, seed, 0, None, 0, 0, _p(out), _p(counts), _p(reads))
    return out, counts, reads


def walk_crafted(hits, K, read_len, from_=0, to=1 << 30, low_coef=0.0, high_coef=1.25):
    """KAT hook: getInitState + computeCount (the literal updateCount state machine) over a crafted LongHit list.
    Returns int32[n, 5] = (start, end, count, startIndex, endIndex) after every step."""
    h = _i32(hits).reshape(-1, 3)
    out = np.zeros((len(h), 5), np.int32)
    L = lib()
    L.vo_walk_crafted.argtypes = [C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.vo_walk_crafted.restype = None
    L.vo_walk_crafted(K, low_coef, high_coef, from_, to, read_len, _p(h), len(h), _p(out))
    return out


def pair_select(forward1, reverse1, forward2, reverse2, fragment_ends=(1 << 30,), is_fragmented=False, insert_len=1000):
    """KAT hook: the decision of Mapper.mapRead (Mapper.java:100-389) on crafted candidate lists of (start, end, count).
    Returns dict(cls, r1=(side, index) or None, r2=..., frag1, frag2)."""
    fe = _i32(fragment_ends)
    lists = [_i32(x).reshape(-1, 3) for x in (forward1, reverse1, forward2, reverse2)]
    out = np.zeros(7, np.int32)
    L = lib()
    L.vo_pair_select.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p, C.c_int] * 4 + [C.c_void_p]
    L.vo_pair_select.restype = None
    args = []
    for a in lists:
        args += [_p(a if len(a) else np.zeros((1, 3), np.int32)), len(a)]
    L.vo_pair_select(_p(fe), len(fe), int(is_fragmented), insert_len, *args, _p(out))
    o = [int(x) for x in out]
    return dict(cls=o[0], r1=(o[1], o[2]) if o[2] >= 0 else None, r2=(o[3], o[4]) if o[4] >= 0 else None, frag1=o[5], frag2=o[6])
