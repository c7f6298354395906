"""Host-side mirror of the reference's batch API for the read-mapping hot path.

The reference is Java and no JVM exists in this environment, so the host layer above the C ABI is
written in Python with the same names, argument meaning and error behaviour as the Java classes:

    Mapper.mapReads(reads, genome)                  Mapper.java:420
    Mapper.mapReads(mappedData, genome)             Mapper.java:477   (remap of needToRemap)
    MapperMSA.mapAllReads(reads, refAlignment)      MapperMSA.java:274
    ConsensusBuilderSimple.getConsensusSimple(...)  ConsensusBuilderSimple.java:123

Everything here is plumbing (ctypes + numpy); all computation happens in libvirgena_gpu.so.
The Java shim a maintainer would add instead is in virgena_b200/java/ and INTEGRATION.md.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import BAM_DTYPE, PAIR_DTYPE, VgError, VgMapStats, VgParams, ptr

NULL_SEQ = b"*"  # Constants.NULL_SEQ (Constants.java:8)


def _u8(x):
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8)
    if isinstance(x, str):
        return np.frombuffer(x.encode("latin1"), dtype=np.uint8)
    return np.ascontiguousarray(x, dtype=np.uint8)


class Reference:
    """Reference (Reference.java): concatenated sequence, contig/fragment ends and the thread count that
    StringIndexer.buildIndexPara would use (it determines the duplicated boundary positions, Q1)."""

    def __init__(self, seq, K=5, contig_ends=None, fragment_ends=None, is_fragmented=False, thread_num=1,
                 index_csr=None):
        self.seqB = _u8(seq)
        self.length = len(self.seqB)
        self.K = K
        self.contigEnds = np.ascontiguousarray(contig_ends if contig_ends is not None else [self.length], np.int32)
        self.fragmentEnds = np.ascontiguousarray(fragment_ends if fragment_ends is not None else self.contigEnds,
                                                 np.int32)
        self.isFragmented = bool(is_fragmented)
        self.threadNum = int(thread_num)
        self.index_csr = index_csr  # optional (keys[nk,K] u8, offsets[nk+1] i64, positions i32) built by the host

    def derive(self, seq):
        """new Reference(byte[] g, K, genome, threadNum) (Reference.java:185-205): same layout, new bases."""
        return Reference(seq, self.K, self.contigEnds, self.fragmentEnds, self.isFragmented, self.threadNum)


class ReferenceAlignment:
    """ReferenceAlignment (ReferenceAlignment.java): gapped rows of equal length and the index
    k-mer -> sorted distinct alignment columns (buildAlnIndex :138-209), built on the host."""

    def __init__(self, rows, K=7):
        rows = np.ascontiguousarray(rows, np.uint8)
        self.rows = rows
        self.length = rows.shape[1]
        self.refNum = rows.shape[0]
        self.K = K
        table = {}
        for r in range(rows.shape[0]):
            cols = np.nonzero(rows[r] != ord("-"))[0]
            seq = rows[r][cols].tobytes()
            for i in range(len(seq) - K + 1):
                table.setdefault(seq[i:i + K], set()).add(int(cols[i]))
        keys = sorted(table)
        self.keys = np.frombuffer(b"".join(keys), np.uint8).reshape(len(keys), K).copy() if keys else np.zeros((0, K), np.uint8)
        self.offsets = np.zeros(len(keys) + 1, np.int64)
        cols_all = []
        for i, k in enumerate(keys):
            c = sorted(table[k])
            cols_all.extend(c)
            self.offsets[i + 1] = self.offsets[i] + len(c)
        self.columns = np.array(cols_all, np.int32)


class PairedReads:
    """DataReader.pairedReads in the layout of vg_upload_reads: one byte pool + per-mate offsets/lengths."""

    def __init__(self, bases, off1, len1, off2, len2):
        self.bases = _u8(bases)
        self.off1 = np.ascontiguousarray(off1, np.int64)
        self.len1 = np.ascontiguousarray(len1, np.int32)
        self.off2 = np.ascontiguousarray(off2, np.int64)
        self.len2 = np.ascontiguousarray(len2, np.int32)

    def __len__(self):
        return len(self.off1)

    def seq1(self, i):
        return self.bases[self.off1[i]:self.off1[i] + self.len1[i]].tobytes()

    def seq2(self, i):
        return self.bases[self.off2[i]:self.off2[i] + self.len2[i]].tobytes()

    def packed(self):
        """(packed, exc_pos, exc_byte): the transfer format of vg_upload_reads_packed (2 bits per ACGT base + exceptions),
        made by the library's host-side packer vg_pack_reads."""
        if getattr(self, "_packed", None) is None:
            L = _lib.lib()
            n = len(self.bases)
            pk = np.zeros((n + 3) // 4 + 16, np.uint8)
            ne = C.c_int64()
            cap = max(1024, n // 64)
            while True:
                pos = np.zeros(cap, np.int64)
                byt = np.zeros(cap, np.uint8)
                rc = L.vg_pack_reads(ptr(self.bases), n, ptr(pk), ptr(pos), ptr(byt), cap, C.byref(ne))
                if rc == 0:
                    break
                if rc != -4:
                    raise VgError(rc, "vg_pack_reads failed")
                cap = ne.value
            self._packed = (pk, pos[:ne.value].copy(), byt[:ne.value].copy())
        return self._packed


class MappedData:
    """MappedData (MappedData.java): per-pair records in input order + the sequence1 byte pool.
    The six Java lists are views by class (task order == input order, Mapper.java:446-461)."""

    def __init__(self, pair_index, records, pool, stats, msa=False, ops=None, reads=None):
        self.pair_index = pair_index      # index of each record's pair in the resident read set
        self.records = records            # PAIR_DTYPE
        self.seq1_pool = pool
        self.stats = stats
        self.msa = msa
        self.needToRemap = np.zeros(0, np.int64)
        self.ops = ops                    # compact results (vg_get_results_compact): edit operations instead of sequence1
        self.reads = reads                # the PairedReads the compact records refer to

    def _cls(self, *classes):
        return np.nonzero(np.isin(self.records["cls"], classes))[0]

    @property
    def concordant(self):
        return self._cls(_lib.CLS_CONCORDANT_FR, _lib.CLS_CONCORDANT_RF)

    @property
    def discordant(self):
        return self._cls(_lib.CLS_DISCORDANT)

    @property
    def leftMateMapped(self):
        return self._cls(_lib.CLS_LEFT)

    @property
    def rightMateMapped(self):
        return self._cls(_lib.CLS_RIGHT)

    @property
    def unmapped(self):
        return self._cls(_lib.CLS_UNMAPPED)

    @property
    def mappedReads(self):
        """(record, mate) pairs in the order of MappedData.mappedReads (r1 then r2 of each pair)."""
        pres = self.records["m"]["present"]
        rec, mate = np.nonzero(pres)
        return rec, mate

    def sequence1(self, rec, mate):
        m = self.records["m"][rec, mate]
        if self.ops is not None:
            p = int(self.pair_index[rec])
            read = self.reads.seq1(p) if mate == 0 else self.reads.seq2(p)
            return sequence1_from_ops(read, bool(m["reverse"]), int(m["start1"]),
                                      self.ops[m["seq1_offset"]:m["seq1_offset"] + m["pad_"]])
        return self.seq1_pool[m["seq1_offset"]:m["seq1_offset"] + m["length"]].tobytes()


_COMPLEMENT = np.zeros(256, np.uint8)   # Constants.complement (Constants.java:11-20): anything unlisted -> 0
for _a, _b in (("A", "T"), ("T", "A"), ("G", "C"), ("C", "G"), ("-", "-"), ("N", "N")):
    _COMPLEMENT[ord(_a)] = ord(_b)
_LOWER = np.zeros(256, np.uint8)        # Constants.lower (Constants.java:22-30)
for _a in "ATGCN":
    _LOWER[ord(_a)] = ord(_a.lower())


def sequence1_from_ops(read, reverse, start1, ops):
    """Alignment.sequence1 from the compact result form: the read in the mate's orientation (reverse =
    Constants.getComplement, i.e. the reverse complement) from start1 on; op 0 copies read bases, op 1 lower-cases them
    (insertion), op 2 writes '-' (deletion)."""
    r = np.frombuffer(read, np.uint8)
    if reverse:
        r = _COMPLEMENT[r[::-1]]
    out = []
    i = start1
    for e in ops:
        ln, op = int(e) >> 4, int(e) & 15
        if op == 0:
            out.append(r[i:i + ln])
            i += ln
        elif op == 1:
            out.append(_LOWER[r[i:i + ln]])
            i += ln
        else:
            out.append(np.full(ln, ord("-"), np.uint8))
    return np.concatenate(out).tobytes() if out else b""


class Context:
    """One vg_ctx: Mapper + KMerCounter + SmithWatermanGotoh parameters bound to a CUDA device."""

    def __init__(self, K=5, low_coef=0.0, high_coef=1.25, cutoffs=None, insert_len=1000, match=2, mismatch=-3,
                 gop=5, gep=2, device=0):
        self.L = _lib.lib()
        self.cutoffs = np.ascontiguousarray(cutoffs if cutoffs is not None else np.zeros(1024, np.int32), np.int32)
        self.K = K
        p = VgParams(K, low_coef, high_coef, self.cutoffs.ctypes.data, len(self.cutoffs), insert_len, match, mismatch,
                     gop, gep)
        h = C.c_void_p()
        rc = self.L.vg_ctx_create(C.byref(p), device, C.byref(h))
        self.h = h
        if rc != 0:
            msg = self.L.vg_last_error(h).decode() if h else "context allocation failed"
            if h:
                self.L.vg_ctx_destroy(h)
                self.h = None
            raise VgError(rc, msg)
        self.n_pairs = 0
        self.G = 0
        self.sw_params = (match, mismatch, gop, gep)

    def close(self):
        if getattr(self, "h", None):
            self.L.vg_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise VgError(rc, self.L.vg_last_error(self.h).decode())

    # --- reference / index -------------------------------------------------------------------
    def set_reference(self, ref: Reference):
        keys = offs = pos = None
        nk = 0
        if ref.index_csr is not None:
            keys, offs, pos = ref.index_csr
            keys = np.ascontiguousarray(keys, np.uint8)
            offs = np.ascontiguousarray(offs, np.int64)
            pos = np.ascontiguousarray(pos, np.int32)
            nk = len(offs) - 1
        self._chk(self.L.vg_set_reference(self.h, ptr(ref.seqB), ref.length, ptr(ref.contigEnds), len(ref.contigEnds),
                                          ptr(ref.fragmentEnds), len(ref.fragmentEnds), int(ref.isFragmented),
                                          ptr(keys), ptr(offs), ptr(pos), nk, ref.threadNum))
        self.G = ref.length

    def set_msa(self, aln: ReferenceAlignment, low_coef, high_coef, cutoffs):
        cut = np.ascontiguousarray(cutoffs, np.int32)
        self._chk(self.L.vg_set_msa_index(self.h, aln.K, low_coef, high_coef, ptr(cut), len(cut), ptr(aln.keys),
                                          ptr(aln.offsets), ptr(aln.columns), len(aln.offsets) - 1, aln.length))

    # --- reads -------------------------------------------------------------------------------
    def upload_reads(self, reads: PairedReads):
        self._chk(self.L.vg_upload_reads(self.h, ptr(reads.bases), len(reads.bases), ptr(reads.off1), ptr(reads.len1),
                                         ptr(reads.off2), ptr(reads.len2), len(reads)))
        self.n_pairs = len(reads)

    def upload_reads_packed(self, reads: PairedReads):
        """vg_upload_reads_packed: the same resident reads, 4x fewer host-to-device bytes."""
        pk, pos, byt = reads.packed()
        self._chk(self.L.vg_upload_reads_packed(self.h, ptr(pk), len(reads.bases), ptr(pos), ptr(byt), len(pos), ptr(reads.off1),
                                                ptr(reads.len1), ptr(reads.off2), ptr(reads.len2), len(reads)))
        self.n_pairs = len(reads)

    def get_results_compact(self):
        """vg_get_results_compact: (records, ops) — seq1_offset / pad_ of a mate address its run-length edit operations."""
        n, _ = self.result_sizes()
        rec = np.zeros(n, PAIR_DTYPE)
        no = C.c_int64()
        cap = max(1024, n * 16)
        while True:
            ops = np.zeros(cap, np.uint32)
            rc = self.L.vg_get_results_compact(self.h, ptr(rec), n, ptr(ops), cap, C.byref(no))
            if rc != -4:
                break
            cap = no.value
        self._chk(rc)
        return rec, ops[:no.value]

    def get_results_compact_into(self, rec, ops):
        """vg_get_results_compact into caller-owned (e.g. pinned) arrays; returns the number of operations."""
        no = C.c_int64()
        self._chk(self.L.vg_get_results_compact(self.h, ptr(rec), len(rec), ptr(ops), len(ops), C.byref(no)))
        return no.value

    def upload_reads_packed_raw(self, packed_ptr, n_bases, exc_pos, exc_byte, off1, len1, off2, len2):
        self._chk(self.L.vg_upload_reads_packed(self.h, C.c_void_p(packed_ptr), n_bases, ptr(exc_pos), ptr(exc_byte), len(exc_pos),
                                                ptr(off1), ptr(len1), ptr(off2), ptr(len2), len(off1)))
        self.n_pairs = len(off1)

    # --- mapping -----------------------------------------------------------------------------
    def map_pairs(self, pair_idx=None, fetch=True):
        st = VgMapStats()
        if pair_idx is None:
            self._chk(self.L.vg_map_pairs(self.h, None, 0, C.byref(st)))
            idx = np.arange(self.n_pairs, dtype=np.int64)
        else:
            idx = np.ascontiguousarray(pair_idx, np.int64)
            self._chk(self.L.vg_map_pairs(self.h, ptr(idx), len(idx), C.byref(st)))
        if not fetch:
            return st.as_array()
        rec, pool = self.get_results()
        return MappedData(idx, rec, pool, st.as_array())

    def map_pairs_msa(self, fetch=True):
        st = VgMapStats()
        self._chk(self.L.vg_map_pairs_msa(self.h, C.byref(st)))
        if not fetch:
            return st.as_array()
        rec, pool = self.get_results()
        return MappedData(np.arange(self.n_pairs, dtype=np.int64), rec, pool, st.as_array(), msa=True)

    def result_sizes(self):
        n = C.c_int64()
        nb = C.c_int64()
        self._chk(self.L.vg_result_sizes(self.h, C.byref(n), C.byref(nb)))
        return n.value, nb.value

    def get_results(self):
        n = C.c_int64()
        nb = C.c_int64()
        self._chk(self.L.vg_result_sizes(self.h, C.byref(n), C.byref(nb)))
        rec = np.zeros(n.value, PAIR_DTYPE)
        pool = np.zeros(max(nb.value, 1), np.uint8)
        self._chk(self.L.vg_get_results(self.h, ptr(rec), n.value, ptr(pool), len(pool)))
        return rec, pool[:nb.value]

    # --- pileup / consensus ------------------------------------------------------------------
    def pileup(self, accumulate=False):
        self._chk(self.L.vg_pileup(self.h, int(accumulate)))

    def counts_device(self):
        p = C.c_void_p()
        n = C.c_int64()
        self._chk(self.L.vg_pileup_counts_device(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def counts(self):
        out = np.zeros((self.G, 5), np.int32)
        self._chk(self.L.vg_pileup_counts_host(self.h, ptr(out), out.size))
        return out

    def consensus(self, genome, save_uncovered=False, coverage_threshold=0):
        genome = _u8(genome)
        out = np.zeros(len(genome), np.uint8)
        self._chk(self.L.vg_consensus(self.h, int(save_uncovered), coverage_threshold, ptr(genome), len(genome),
                                      ptr(out)))
        return out

    # --- batch kernels exposed for the "next" callers and the tests ---------------------------
    def sw_align_batch(self, s1_pool, s1_off, s2_pool, s2_off):
        s1_pool, s2_pool = _u8(s1_pool), _u8(s2_pool)
        s1_off = np.ascontiguousarray(s1_off, np.int64)
        s2_off = np.ascontiguousarray(s2_off, np.int64)
        n = len(s1_off) - 1
        rec = np.zeros((n, 7), np.int32)
        cap = np.diff(s1_off) + np.diff(s2_off)
        soff = np.zeros(n + 1, np.int64)
        np.cumsum(cap, out=soff[1:])
        pool = np.zeros(int(soff[-1]) + 1, np.uint8)
        self._chk(self.L.vg_sw_align_batch(self.h, ptr(s1_pool), ptr(s1_off), ptr(s2_pool), ptr(s2_off), n, ptr(rec),
                                           ptr(pool), ptr(soff)))
        return rec, pool, soff

    def sw_max_score_batch(self, s1_pool, s1_off, s2_pool, s2_off):
        s1_pool, s2_pool = _u8(s1_pool), _u8(s2_pool)
        s1_off = np.ascontiguousarray(s1_off, np.int64)
        s2_off = np.ascontiguousarray(s2_off, np.int64)
        n = len(s1_off) - 1
        out = np.zeros(n, np.int32)
        self._chk(self.L.vg_sw_max_score_batch(self.h, ptr(s1_pool), ptr(s1_off), ptr(s2_pool), ptr(s2_off), n, ptr(out)))
        return out

    def seed_batch(self, pool, off, msa=False, max_windows=64):
        pool = _u8(pool)
        off = np.ascontiguousarray(off, np.int64)
        n = len(off) - 1
        win = np.zeros((n, max_windows, 3), np.int32)
        nwin = np.zeros(n, np.int32)
        self._chk(self.L.vg_seed_batch(self.h, int(msa), ptr(pool), ptr(off), n, max_windows, ptr(win), ptr(nwin)))
        return win, nwin

    def best_window_batch(self, pool, off, msa=False):
        """vg_best_window_batch: KMerCounter.mapReadToRegion(read, index, 0, length) / computeKMerCount for every read.
        Returns (windows int32[n, 3] = start, end, count; found int32[n])."""
        pool = _u8(pool)
        off = np.ascontiguousarray(off, np.int64)
        n = len(off) - 1
        win = np.zeros((max(n, 1), 3), np.int32)
        found = np.zeros(max(n, 1), np.int32)
        self._chk(self.L.vg_best_window_batch(self.h, int(msa), ptr(pool), ptr(off), n, ptr(win), ptr(found)))
        return win[:n], found[:n]

    def random_model_counts(self, genomes, order, seed, first_draw, read_len, n_reads, msa=False, want_reads=False):
        """vg_random_model_counts: n_reads random reads of read_len bases from the order-`order` Markov model of `genomes`
        (list of byte strings), generated and scored (computeKMerCount) on the device. Returns counts int32[n] (unsorted)
        or (counts, reads uint8[n, read_len])."""
        gs = [_u8(g) for g in genomes]
        pool = np.concatenate(gs) if gs else np.zeros(0, np.uint8)
        off = np.zeros(len(gs) + 1, np.int64)
        off[1:] = np.cumsum([len(g) for g in gs])
        counts = np.zeros(max(n_reads, 1), np.int32)
        reads = np.zeros((max(n_reads, 1), read_len), np.uint8) if want_reads else None
        self._chk(self.L.vg_random_model_counts(self.h, int(msa), ptr(pool), ptr(off), len(gs), int(order), C.c_uint64(int(seed)),
                                                int(first_draw), int(read_len), int(n_reads), ptr(counts),
                                                ptr(reads) if want_reads else None))
        return (counts[:n_reads], reads[:n_reads]) if want_reads else counts[:n_reads]

    def map_to_regions(self, region_pool, region_off, read_pool, read_off, read_region):
        """vg_map_to_regions: mapReadToRegion of read i against region read_region[i] (each region its own index).
        Returns (windows int32[n, 3], found int32[n])."""
        region_pool, read_pool = _u8(region_pool), _u8(read_pool)
        region_off = np.ascontiguousarray(region_off, np.int64)
        read_off = np.ascontiguousarray(read_off, np.int64)
        read_region = np.ascontiguousarray(read_region, np.int32)
        n = len(read_off) - 1
        win = np.zeros((max(n, 1), 3), np.int32)
        found = np.zeros(max(n, 1), np.int32)
        self._chk(self.L.vg_map_to_regions(self.h, ptr(region_pool), ptr(region_off), len(region_off) - 1, ptr(read_pool),
                                           ptr(read_off), n, ptr(read_region), ptr(win), ptr(found)))
        return win[:n], found[:n]

    def pileup_alignments(self, seq1_pool, seq1_off, length, pos, total_positions, end=None):
        """vg_pileup_alignments; end[i] = end of alignment i's own region (None: the whole array)."""
        seq1_pool = _u8(seq1_pool)
        seq1_off = np.ascontiguousarray(seq1_off, np.int64)
        length = np.ascontiguousarray(length, np.int32)
        pos = np.ascontiguousarray(pos, np.int64)
        end = None if end is None else np.ascontiguousarray(end, np.int64)
        counts = np.zeros((max(int(total_positions), 1), 5), np.int32)
        self._chk(self.L.vg_pileup_alignments(self.h, ptr(seq1_pool), ptr(seq1_off), ptr(length), ptr(pos), ptr(end),
                                              len(length), int(total_positions), ptr(counts)))
        return counts[:int(total_positions)]

    def set_results(self, md):
        """vg_set_results: a host MappedData becomes the context's last mapping (pileup / BAM fields / identity use it)."""
        rec = np.ascontiguousarray(md.records, PAIR_DTYPE)
        idx = np.ascontiguousarray(md.pair_index, np.int64)
        pool = _u8(md.seq1_pool)
        self._chk(self.L.vg_set_results(self.h, ptr(rec), len(rec), ptr(idx), ptr(pool) if len(pool) else None, len(pool)))

    def identity_counting(self, window_size, thread_num=1):
        """vg_identity_counting: (identity float32[G], coverage int32[G]) of the last mapping."""
        ident = np.zeros(max(self.G, 1), np.float32)
        cov = np.zeros(max(self.G, 1), np.int32)
        self._chk(self.L.vg_identity_counting(self.h, int(window_size), int(thread_num), ptr(ident), ptr(cov)))
        return ident[:self.G], cov[:self.G]

    def debug_seed_hits(self, read, msa=False):
        """Test hook: LongHits (genomePos, readPos, len) of one read after mergeHits + splitLongHits."""
        read = _u8(read)
        fn = self.L.vg_debug_seed_hits
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        out = np.zeros((16384, 3), np.int32)
        n = C.c_int32()
        self._chk(fn(self.h, int(msa), ptr(read), len(read), ptr(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def timings(self):
        t = np.zeros(6, np.float64)
        self._chk(self.L.vg_last_timings(self.h, ptr(t)))
        return dict(seed_ms=t[0], sw_fill_ms=t[1], traceback_ms=t[2], pileup_ms=t[3], total_ms=t[4], sw_cells=t[5])

    def seed_timings(self):
        t = np.zeros(8, np.float64)
        self._chk(self.L.vg_seed_timings(self.h, ptr(t)))
        return dict(gen_ms=t[0], win_ms=t[1], fallback_ms=t[2], longhits_in=t[3], longhits_kept=t[4], fallback_items=t[5],
                    windows_out=t[6], items=t[7])

    def dpx_peak(self):
        v = C.c_double()
        self._chk(self.L.vg_dpx_peak(self.h, C.byref(v)))
        return v.value

    def stream(self):
        p = C.c_void_p()
        self._chk(self.L.vg_ctx_stream(self.h, C.byref(p)))
        return p.value or 0

    def upload_reads_raw(self, bases_ptr, n_bases, off1, len1, off2, len2):
        """vg_upload_reads from a caller-owned (e.g. pinned) host byte pool given as a raw address."""
        self._chk(self.L.vg_upload_reads(self.h, C.c_void_p(bases_ptr), n_bases, ptr(off1), ptr(len1), ptr(off2),
                                         ptr(len2), len(off1)))
        self.n_pairs = len(off1)

    def get_results_into(self, rec, pool):
        """vg_get_results into caller-owned (e.g. pinned) arrays; returns (n_results, seq1_bytes)."""
        n = C.c_int64()
        nb = C.c_int64()
        self._chk(self.L.vg_result_sizes(self.h, C.byref(n), C.byref(nb)))
        self._chk(self.L.vg_get_results(self.h, ptr(rec), len(rec), ptr(pool), len(pool)))
        return n.value, nb.value

    def kernel_launches(self):
        return int(self.L.vg_kernel_launches(self.h))

    # --- multi-GPU, one process per GPU (include/virgena_gpu.h, "Multi-GPU" (a)) ----------------------------
    @staticmethod
    def comm_unique_id():
        """vg_comm_unique_id: 128 bytes that rank 0 hands to the other ranks by any channel the host has."""
        buf = np.zeros(128, np.uint8)
        rc = _lib.lib().vg_comm_unique_id(ptr(buf))
        if rc != 0:
            raise VgError(rc, "NCCL is not available (libnccl.so.2 could not be loaded)")
        return buf

    def comm_init_rank(self, uid, rank, n_ranks):
        uid = np.ascontiguousarray(uid, np.uint8)
        assert len(uid) == 128
        self._chk(self.L.vg_comm_init_rank(self.h, ptr(uid), int(rank), int(n_ranks)))

    def pileup_allreduce(self):
        """vg_pileup_allreduce: the counts of all ranks summed in place (no-op without a communicator)."""
        self._chk(self.L.vg_pileup_allreduce(self.h))

    def allreduce_ms(self):
        v = C.c_double()
        self._chk(self.L.vg_last_allreduce_ms(self.h, C.byref(v)))
        return v.value

    def bam_fields(self):
        """vg_bam_fields: (records BAM_DTYPE[2 * n_results], CIGAR pool uint32) of the last single-reference mapping."""
        n = C.c_int64()
        self._chk(self.L.vg_result_sizes(self.h, C.byref(n), None))
        rec = np.zeros(2 * n.value, BAM_DTYPE)
        nops = C.c_int64()
        pool = np.zeros(max(1, 2 * n.value * 8), np.uint32)
        rc = self.L.vg_bam_fields(self.h, ptr(rec), len(rec), ptr(pool), len(pool), C.byref(nops))
        if rc == -4:  # VG_ERR_CAPACITY: n_ops now holds the size
            pool = np.zeros(max(1, nops.value), np.uint32)
            rc = self.L.vg_bam_fields(self.h, ptr(rec), len(rec), ptr(pool), len(pool), C.byref(nops))
        self._chk(rc)
        return rec, pool[:nops.value]


class _BorrowedContext(Context):
    """A Context whose vg_ctx belongs to somebody else (vg_multi_ctx): never destroyed from here."""

    def __init__(self, owner, handle):
        self.L, self.h, self._owner = owner.L, handle, owner
        self.K, self.G, self.n_pairs, self.sw_params, self.cutoffs = owner.K, owner.G, 0, owner.sw_params, owner.cutoffs

    def close(self):
        self.h = None


class MultiContext:
    """One vg_multi: the same Mapper bound to several GPUs of this process (include/virgena_gpu.h, "Multi-GPU" (c)). Pairs
    are sharded in contiguous blocks, the index is replicated, pileup() all-reduces the counts with NCCL inside the
    library. The call surface is the subset of Context that Mapper / ConsensusBuilderSimple use."""

    def __init__(self, devices, K=5, low_coef=0.0, high_coef=1.25, cutoffs=None, insert_len=1000, match=2, mismatch=-3,
                 gop=5, gep=2):
        self.L = _lib.lib()
        self.cutoffs = np.ascontiguousarray(cutoffs if cutoffs is not None else np.zeros(1024, np.int32), np.int32)
        self.K = K
        self.devices = np.ascontiguousarray(devices, np.int32)
        p = VgParams(K, low_coef, high_coef, self.cutoffs.ctypes.data, len(self.cutoffs), insert_len, match, mismatch,
                     gop, gep)
        h = C.c_void_p()
        rc = self.L.vg_multi_create(C.byref(p), ptr(self.devices), len(self.devices), C.byref(h))
        self.h = h
        if rc != 0:
            msg = self.L.vg_multi_last_error(h).decode() if h else "allocation failed"
            if h:
                self.L.vg_multi_destroy(h)
                self.h = None
            raise VgError(rc, msg)
        self.n_pairs = 0
        self.G = 0
        self.sw_params = (match, mismatch, gop, gep)

    def close(self):
        if getattr(self, "h", None):
            self.L.vg_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise VgError(rc, self.L.vg_multi_last_error(self.h).decode())

    def device_ctx(self, i):
        """Borrowed Context view of the i-th device's vg_ctx (timings, launches, BAM fields of its shard, ...); it stays
        owned by the multi-context."""
        return _BorrowedContext(self, C.c_void_p(self.L.vg_multi_ctx(self.h, i)))

    def set_reference(self, ref: Reference):
        keys = offs = pos = None
        nk = 0
        if ref.index_csr is not None:
            keys, offs, pos = ref.index_csr
            keys = np.ascontiguousarray(keys, np.uint8)
            offs = np.ascontiguousarray(offs, np.int64)
            pos = np.ascontiguousarray(pos, np.int32)
            nk = len(offs) - 1
        self._chk(self.L.vg_multi_set_reference(self.h, ptr(ref.seqB), ref.length, ptr(ref.contigEnds), len(ref.contigEnds),
                                                ptr(ref.fragmentEnds), len(ref.fragmentEnds), int(ref.isFragmented),
                                                ptr(keys), ptr(offs), ptr(pos), nk, ref.threadNum))
        self.G = ref.length

    def set_msa(self, aln: ReferenceAlignment, low_coef, high_coef, cutoffs):
        cut = np.ascontiguousarray(cutoffs, np.int32)
        self._chk(self.L.vg_multi_set_msa_index(self.h, aln.K, low_coef, high_coef, ptr(cut), len(cut), ptr(aln.keys),
                                                ptr(aln.offsets), ptr(aln.columns), len(aln.offsets) - 1, aln.length))

    def upload_reads(self, reads: PairedReads):
        self._chk(self.L.vg_multi_upload_reads(self.h, ptr(reads.bases), len(reads.bases), ptr(reads.off1), ptr(reads.len1),
                                               ptr(reads.off2), ptr(reads.len2), len(reads)))
        self.n_pairs = len(reads)

    def upload_reads_raw(self, bases_ptr, n_bases, off1, len1, off2, len2):
        self._chk(self.L.vg_multi_upload_reads(self.h, C.c_void_p(bases_ptr), n_bases, ptr(off1), ptr(len1), ptr(off2),
                                               ptr(len2), len(off1)))
        self.n_pairs = len(off1)

    def upload_reads_packed(self, reads: PairedReads):
        pk, pos, byt = reads.packed()
        self._chk(self.L.vg_multi_upload_reads_packed(self.h, ptr(pk), len(reads.bases), ptr(pos), ptr(byt), len(pos),
                                                      ptr(reads.off1), ptr(reads.len1), ptr(reads.off2), ptr(reads.len2), len(reads)))
        self.n_pairs = len(reads)

    def get_results_compact(self):
        n = C.c_int64()
        self._chk(self.L.vg_multi_result_sizes(self.h, C.byref(n), None))
        rec = np.zeros(n.value, PAIR_DTYPE)
        no = C.c_int64()
        cap = max(1024, n.value * 16)
        while True:
            ops = np.zeros(cap, np.uint32)
            rc = self.L.vg_multi_get_results_compact(self.h, ptr(rec), n.value, ptr(ops), cap, C.byref(no))
            if rc != -4:
                break
            cap = no.value
        self._chk(rc)
        return rec, ops[:no.value]

    def get_results(self):
        n = C.c_int64()
        nb = C.c_int64()
        self._chk(self.L.vg_multi_result_sizes(self.h, C.byref(n), C.byref(nb)))
        rec = np.zeros(n.value, PAIR_DTYPE)
        pool = np.zeros(max(nb.value, 1), np.uint8)
        self._chk(self.L.vg_multi_get_results(self.h, ptr(rec), n.value, ptr(pool), len(pool)))
        return rec, pool[:nb.value]

    def get_results_into(self, rec, pool):
        n = C.c_int64()
        nb = C.c_int64()
        self._chk(self.L.vg_multi_result_sizes(self.h, C.byref(n), C.byref(nb)))
        self._chk(self.L.vg_multi_get_results(self.h, ptr(rec), len(rec), ptr(pool), len(pool)))
        return n.value, nb.value

    def map_pairs(self, pair_idx=None, fetch=True):
        st = VgMapStats()
        if pair_idx is None:
            self._chk(self.L.vg_multi_map_pairs(self.h, None, 0, C.byref(st)))
            idx = np.arange(self.n_pairs, dtype=np.int64)
        else:
            idx = np.ascontiguousarray(pair_idx, np.int64)
            self._chk(self.L.vg_multi_map_pairs(self.h, ptr(idx), len(idx), C.byref(st)))
        if not fetch:
            return st.as_array()
        rec, pool = self.get_results()
        return MappedData(idx, rec, pool, st.as_array())

    def map_pairs_msa(self, fetch=True):
        st = VgMapStats()
        self._chk(self.L.vg_multi_map_pairs_msa(self.h, C.byref(st)))
        if not fetch:
            return st.as_array()
        rec, pool = self.get_results()
        return MappedData(np.arange(self.n_pairs, dtype=np.int64), rec, pool, st.as_array(), msa=True)

    def pileup(self, accumulate=False):
        """vg_pileup on every device + the NCCL all-reduce of the counts (ConsensusBuilderSimple.java:138-145)."""
        self._chk(self.L.vg_multi_pileup(self.h, int(accumulate)))

    def counts(self):
        out = np.zeros((self.G, 5), np.int32)
        self._chk(self.L.vg_multi_pileup_counts_host(self.h, ptr(out), out.size))
        return out

    def consensus(self, genome, save_uncovered=False, coverage_threshold=0):
        genome = _u8(genome)
        out = np.zeros(len(genome), np.uint8)
        self._chk(self.L.vg_multi_consensus(self.h, int(save_uncovered), coverage_threshold, ptr(genome), len(genome),
                                            ptr(out)))
        return out


CIGAR_OPS = "MID?S"


def cigar_string(ops):
    """CIGAR text of a run of len << 4 | op elements ('*' when empty)."""
    return "".join("%d%s" % (int(o) >> 4, CIGAR_OPS[int(o) & 15]) for o in ops) or "*"


class ContigBuilder:
    """The read-counting step of ContigBuilder (ContigBuilder.java:72-104, rebuilt at :219-227): every read of a cluster is
    seeded against the cluster's centroid (mapReadToRegion with the centroid's own index), aligned against the window
    and piled up into the centroid's counts. One call handles all clusters."""

    def __init__(self, mapper):
        self.ctx = mapper.ctx

    def countReads(self, centroids, reads, read_cluster):
        """centroids, reads: lists of byte strings; read_cluster[i] = cluster of read i. Returns (windows int32[n, 3],
        found int32[n], alignments int32[n, 7], sequence1 list, counts per centroid list of int32[len, 5])."""
        cpool = np.frombuffer(b"".join(bytes(c) for c in centroids), np.uint8) if centroids else np.zeros(0, np.uint8)
        coff = np.zeros(len(centroids) + 1, np.int64)
        np.cumsum([len(c) for c in centroids], out=coff[1:])
        rpool = np.frombuffer(b"".join(bytes(r) for r in reads), np.uint8) if reads else np.zeros(0, np.uint8)
        roff = np.zeros(len(reads) + 1, np.int64)
        np.cumsum([len(r) for r in reads], out=roff[1:])
        read_cluster = np.ascontiguousarray(read_cluster, np.int32)
        win, found = self.ctx.map_to_regions(cpool, coff, rpool, roff, read_cluster)
        idx = np.nonzero(found)[0]
        # aligner.align(read.seq, centroid[start:end]) (ContigBuilder.java:85-89)
        s2 = [bytes(centroids[read_cluster[i]][win[i, 0]:win[i, 1]]) for i in idx]
        s1 = [bytes(reads[i]) for i in idx]
        s1off = np.zeros(len(idx) + 1, np.int64)
        np.cumsum([len(x) for x in s1], out=s1off[1:])
        s2off = np.zeros(len(idx) + 1, np.int64)
        np.cumsum([len(x) for x in s2], out=s2off[1:])
        aln = np.zeros((len(reads), 7), np.int32)
        seqs = [b""] * len(reads)
        total = int(coff[-1])
        if len(idx):
            rec, pool, soff = self.ctx.sw_align_batch(b"".join(s1), s1off, b"".join(s2), s2off)
            aln[idx] = rec
            for t, i in enumerate(idx):
                seqs[i] = pool[soff[t]:soff[t] + rec[t, 6]].tobytes()
            pos = coff[read_cluster[idx]] + win[idx, 0] + rec[:, 3]  # centroid base + mRead.start + alignment.start2
            counts = self.ctx.pileup_alignments(pool, soff[:-1], rec[:, 6], pos, total, end=coff[read_cluster[idx] + 1])
        else:
            counts = np.zeros((total, 5), np.int32)
        return win, found, aln, seqs, [counts[coff[c]:coff[c + 1]] for c in range(len(centroids))]


class Graph:
    """The reference-scoring step of Graph (Graph.java:493-586): every cluster contig (gaps removed) is seeded
    against each candidate reference sub-sequence with that sub-sequence's own index (mapReadToRegion), aligned against
    the window, and scored; the best reference and those within `delta` of it are kept. One call handles all clusters."""

    def __init__(self, mapper, delta=0.0):
        self.ctx = mapper.ctx
        self.match = mapper.ctx.sw_params[0]
        self.delta = np.float32(delta)

    def SWScore(self, contigs, ref_subseqs):
        """contigs[c]: gapped contig bytes; ref_subseqs[c]: list of byte strings (reference.seq.substring(ref.start, ref.end)
        for every RefSubseq of the cluster's window). Returns one dict per cluster: markedForDeletion, score, identity,
        coverage (float32, as SWScore's fields), best (index of the best reference or -1), refs [(index, score)] kept, and
        per reference: found, window, alignment (7 ints), scores (score, identity, coverage as float32)."""
        seqs = [bytes(b for b in bytes(c) if b != ord("-")) for c in contigs]
        regions, reads, read_region, owner = [], [], [], []
        for c, subs in enumerate(ref_subseqs):
            for j, sub in enumerate(subs):
                read_region.append(len(regions))
                regions.append(bytes(sub))
                reads.append(seqs[c])
                owner.append((c, j))
        n = len(reads)
        goff = np.zeros(len(regions) + 1, np.int64)
        np.cumsum([len(x) for x in regions], out=goff[1:])
        roff = np.zeros(n + 1, np.int64)
        np.cumsum([len(x) for x in reads], out=roff[1:])
        win = np.zeros((max(n, 1), 3), np.int32)
        found = np.zeros(max(n, 1), np.int32)
        if n:
            win, found = self.ctx.map_to_regions(np.frombuffer(b"".join(regions), np.uint8), goff,
                                                 np.frombuffer(b"".join(reads), np.uint8), roff,
                                                 np.asarray(read_region, np.int32))
        idx = [i for i in range(n) if found[i]]
        aln = np.zeros((max(n, 1), 7), np.int32)
        if idx:
            s1 = [reads[i] for i in idx]
            s2 = [regions[i][win[i, 0]:win[i, 1]] for i in idx]
            s1off = np.zeros(len(idx) + 1, np.int64)
            np.cumsum([len(x) for x in s1], out=s1off[1:])
            s2off = np.zeros(len(idx) + 1, np.int64)
            np.cumsum([len(x) for x in s2], out=s2off[1:])
            rec, _, _ = self.ctx.sw_align_batch(b"".join(s1), s1off, b"".join(s2), s2off)
            aln[idx] = rec
        out = []
        k = 0
        for c, subs in enumerate(ref_subseqs):
            ln = len(seqs[c])
            full = ln * self.match
            per, scored = [], []
            coverage = np.float32(0)
            for j in range(len(subs)):
                i = k + j
                entry = dict(found=bool(found[i]), window=tuple(int(x) for x in win[i]), alignment=aln[i].copy(), scores=None)
                if found[i]:
                    with np.errstate(divide="ignore", invalid="ignore"):
                        sc = np.float32(aln[i, 0]) / np.float32(full)
                        ident = np.float32(aln[i, 5]) / np.float32(aln[i, 6])
                        coverage = np.float32(aln[i, 2] - aln[i, 1]) / np.float32(ln)  # RefScore's ctor writes SWScore.coverage
                    entry["scores"] = (sc, ident, coverage)
                    scored.append((j, sc, ident))
                per.append(entry)
            k += len(subs)
            res = dict(markedForDeletion=not scored, score=np.float32(0), identity=np.float32(0), coverage=coverage, best=-1,
                       refs=[], per_ref=per)
            if scored:
                # Collections.sort by descending score, stable (Graph.java:521-527, :562)
                order = sorted(range(len(scored)), key=lambda t: -float(scored[t][1]))
                bj, bs, bi = scored[order[0]]
                if bs > 0:
                    res["refs"] = [(scored[t][0], scored[t][1]) for t in order if np.float32(bs - scored[t][1]) <= self.delta]
                    res["score"], res["identity"], res["best"] = bs, bi, bj
            out.append(res)
        return out


class ReferenceFinder:
    """ReferenceFinder.SWScoreCounter (ReferenceFinder.java:291-322): clusterIDToScores[v][j] = align(vertex contig,
    reference j's sub-sequence under the contig's alignment columns).score; 0 where the reference has no such columns."""

    def __init__(self, mapper):
        self.ctx = mapper.ctx

    def SWScoreCounter(self, vertex_seqs, ref_windows):
        """ref_windows[v][j]: bytes (ref.seq.substring(start, end)) or None (alnToSeqStart/End gave -1). Returns int32[v, j]."""
        nref = max((len(r) for r in ref_windows), default=0)
        scores = np.zeros((len(vertex_seqs), nref), np.int32)
        s1, s2, where = [], [], []
        for v, refs in enumerate(ref_windows):
            for j, w in enumerate(refs):
                if w is not None:
                    s1.append(bytes(vertex_seqs[v]))
                    s2.append(bytes(w))
                    where.append((v, j))
        if where:
            s1off = np.zeros(len(s1) + 1, np.int64)
            np.cumsum([len(x) for x in s1], out=s1off[1:])
            s2off = np.zeros(len(s2) + 1, np.int64)
            np.cumsum([len(x) for x in s2], out=s2off[1:])
            rec, _, _ = self.ctx.sw_align_batch(b"".join(s1), s1off, b"".join(s2), s2off)
            for (v, j), r in zip(where, rec):
                scores[v, j] = r[0]
        return scores


class RandomModel:
    """RandomModel / RandomModelMSA (RandomModel.java:25-145, RandomModelMSA.java:56-135) + KMerCounterBase.computeCuttofs
    (:229-251): the null distribution of the k-mer score of random reads and the cutoffs derived from it. Reads are generated
    (MarkovModel.generate) and scored (computeKMerCount) on the device; sorting and the quantile stay on the host.

    The model's counter is its own KMerCounter instance with lowCoef = 1 / coef (KMerCounter.java:51-55, Q3), hence its own
    context. The reference's RNG is unseeded (Q10); here the reads are those of one SplitMix64(seed) stream consumed length by
    length from maxReadLen down, read by read, which is also what the oracle's seeded restatement draws."""

    def __init__(self, genome, K=5, order=4, coef=1.25, seed=1, device=0):
        self.K, self.modelK, self.coef, self.seed = K, order, float(coef), int(seed)
        self.msa = isinstance(genome, ReferenceAlignment)
        self.ctx = Context(K, float(np.float32(1.0) / np.float32(coef)), float(coef), np.zeros(1, np.int32), device=device)
        if self.msa:
            self.ctx.set_msa(genome, float(np.float32(1.0) / np.float32(coef)), float(coef), np.zeros(1, np.int32))
            self.genomes = [row[row != ord("-")] for row in genome.rows]          # RandomModelMSA.java:63-66
        else:
            self.ctx.set_reference(genome)
            if genome.isFragmented:                                                 # RandomModel.java:88-94
                ends = [0] + [int(e) for e in genome.contigEnds]
                self.genomes = [genome.seqB[a:b] for a, b in zip(ends[:-1], ends[1:])]
            else:
                self.genomes = [genome.seqB]
        self.counts = None

    def genModel(self, readNum, minReadLen, maxReadLen, step):
        """counts[size][readNum], every row sorted (genRandomScores ends with Arrays.sort)."""
        size = (maxReadLen - minReadLen) // step + 1
        self.counts = np.zeros((size, readNum), np.int32)
        draw = 0
        ln = maxReadLen
        for i in range(size - 1, -1, -1):
            c = self.ctx.random_model_counts(self.genomes, self.modelK, self.seed, draw, ln, readNum, msa=self.msa)
            self.counts[i] = np.sort(c)
            draw += readNum * (2 + ln - self.modelK)
            ln -= step
        return self.counts

    @staticmethod
    def computeCuttofs(minReadLen, maxReadLen, step, counts, pValue):
        """KMerCounterBase.computeCuttofs (:229-251) verbatim, including `num = counts.length` (the number of read-length
        bins, not readNum) as the base of the quantile index."""
        num = len(counts)
        size = (maxReadLen - minReadLen) // step + 1
        index = int(np.floor(np.float32(num) * (np.float32(1.0) - np.float32(pValue)) + np.float32(0.5)))  # Math.round
        interp = [int(counts[i][index]) for i in range(size)]
        cutoffs = np.zeros(maxReadLen + 1, np.int32)
        cutoffs[:minReadLen + 1] = interp[0]
        for i in range(maxReadLen, minReadLen, -1):
            right = size - (maxReadLen - i) // step - 1
            if right == 0:
                cutoffs[i] = interp[0]
            else:
                d_right = (maxReadLen - i) % step
                d_left = step - d_right
                v = d_right * interp[right - 1] + d_left * interp[right]
                cutoffs[i] = int(v / step) if v >= 0 else -int(-v / step)  # Java int division truncates
        return cutoffs

    def cutoffs(self, readNum=10000, minReadLen=250, maxReadLen=250, step=10, pValue=0.01):
        self.genModel(readNum, minReadLen, maxReadLen, step)
        return self.computeCuttofs(minReadLen, maxReadLen, step, self.counts, pValue)


class ProblemIntervalBuilder:
    """The identity / coverage profile of ProblemIntervalBuilder (ProblemIntervalBuilder.java:35-160) over the last mapping
    of `mapper`: sliding windows of `window_size` read bases, binary32 sums in the order fixed by `thread_num`."""

    def __init__(self, mapper, window_size, thread_num=1):
        self.mapper, self.windowSize, self.threadNum = mapper, window_size, thread_num

    def countIdentity(self):
        return self.mapper.ctx.identity_counting(self.windowSize, self.threadNum)


class BamPrinter:
    """The per-record derivations of BamPrinter (BamPrinter.java:16-171) for the last mapping of `mapper`: flags,
    positions inside the contig, mate fields, CIGAR and XN, computed on the device. Serialisation stays with the host."""

    def __init__(self, mapper):
        self.mapper = mapper

    def records(self):
        return self.mapper.ctx.bam_fields()


class Mapper:
    """Mapper (Mapper.java). `counter` carries K / cutoffs / coefficients like KMerCounter; the aligner
    parameters are those of SmithWatermanGotoh. One Mapper owns one device context; reads are uploaded once
    and stay resident across mapping passes (iterations of ConsensusBuilderWithReassembling)."""

    def __init__(self, K=5, pValue=0.01, indel_tolerance=1.25, low_coef=0.0, cutoffs=None, insert_len=1000,
                 match=2, mismatch=-3, gop=5, gep=2, device=0, devices=None):
        if devices is not None:  # several GPUs of this process: sharded pairs + NCCL all-reduce inside the library
            self.ctx = MultiContext(devices, K, low_coef, indel_tolerance, cutoffs, insert_len, match, mismatch, gop, gep)
        else:
            self.ctx = Context(K, low_coef, indel_tolerance, cutoffs, insert_len, match, mismatch, gop, gep, device)
        self._reads = None
        self._genome = None

    def _bind(self, reads, genome):
        if reads is not None and reads is not self._reads:
            self.ctx.upload_reads(reads)
            self._reads = reads
        if genome is not self._genome:
            self.ctx.set_reference(genome)
            self._genome = genome

    def mapReads(self, reads_or_mapped, genome):
        """mapReads(ArrayList<PairedRead>, Reference) -> MappedData (Mapper.java:420-475), or
        mapReads(MappedData, Reference) (Mapper.java:477-536): remaps the pairs of mappedData.needToRemap (row indices of
        md.records) and appends their results to the lists the caller kept — md.records / pair_index / seq1_pool become
        [kept rows in their order, remapped rows in needToRemap order], md.stats the figures of :479-523 — and installs the
        union as the context's last mapping, so that getConsensusSimple, BamPrinter and ProblemIntervalBuilder see
        mappedData.mappedReads = kept + remapped like the Java."""
        if isinstance(reads_or_mapped, MappedData):
            md = reads_or_mapped
            self._bind(None, genome)
            need = np.ascontiguousarray(md.needToRemap, np.int64)
            keep = np.setdiff1d(np.arange(len(md.records), dtype=np.int64), need, assume_unique=False)
            kept = md.records[keep]
            # stats of Mapper.java:479-497 over what the caller kept
            kp = kept["m"]["present"] == 1
            conc = np.isin(kept["cls"], (_lib.CLS_CONCORDANT_FR, _lib.CLS_CONCORDANT_RF))
            n_conc = int(conc.sum())
            n_kept_mapped = int(kp.sum())
            st = np.zeros(9, np.int64)
            st[0] = len(need) + n_conc + int((kept["cls"] == _lib.CLS_LEFT).sum()) + int((kept["cls"] == _lib.CLS_RIGHT).sum())
            st[1] = n_conc                                        # bothExist starts from concordant.size()
            st[2] = n_kept_mapped                                 # readsTotal starts from mappedReads.size()
            st[3] = n_kept_mapped
            st[4] = int((kp & (kept["m"]["reverse"] == 0)).sum())
            st[5] = int((kp & (kept["m"]["reverse"] != 0)).sum())
            st[6] = n_conc
            st[7] = int(kept["m"]["score"][kp].sum())
            st[8] = n_conc
            if genome.length == 0 or len(need) == 0:
                # (the Java would still run an empty loop; an empty reference cannot be handed to this overload upstream:
                # remapReadsToAssembly returns before it, ConsensusBuilderWithReassembling.java:742-754)
                sub_rec = np.zeros(0, PAIR_DTYPE)
                sub_pool = np.zeros(0, np.uint8)
                sub_idx = np.zeros(0, np.int64)
            else:
                sub = self.ctx.map_pairs(md.pair_index[need])
                sub_rec, sub_pool, sub_idx = sub.records, sub.seq1_pool, sub.pair_index
                st[1:] += sub.stats[1:]                           # :508-523 (task sums; totalPairs is not touched)
            # kept sequence1 bytes are compacted in front of the remapped ones
            pools, recs = [], kept.copy()
            at = 0
            for r in range(len(recs)):
                for k in range(2):
                    m = recs["m"][r, k]
                    if m["present"] and m["seq1_offset"] >= 0:
                        o, ln = int(m["seq1_offset"]), int(m["length"])
                        pools.append(md.seq1_pool[o:o + ln])
                        recs["m"]["seq1_offset"][r, k] = at
                        at += ln
            sub_rec = sub_rec.copy()
            pres = (sub_rec["m"]["present"] == 1) & (sub_rec["m"]["seq1_offset"] >= 0)
            sub_rec["m"]["seq1_offset"][pres] += at
            md.records = np.concatenate([recs, sub_rec])
            md.pair_index = np.concatenate([md.pair_index[keep], sub_idx])
            md.seq1_pool = np.concatenate(pools + [sub_pool]) if (pools or len(sub_pool)) else np.zeros(0, np.uint8)
            md.stats = st
            md.needToRemap = np.zeros(0, np.int64)
            self.ctx.set_results(md)
            return None
        reads = reads_or_mapped
        self._bind(reads, genome)
        md = self.ctx.map_pairs()  # an empty reference leaves every pair unmapped inside the library (Mapper.java:422-426)
        if genome.length == 0:
            md.stats = np.zeros(9, np.int64)  # the Java returns before any statistic is taken
        return md


def mark_remap_reads(md, problem_intervals, genome_to_assembly, reads):
    """ConsensusBuilderWithReassembling.markRemapReads + adjustCoord (ConsensusBuilderWithReassembling.java:662-739) on a
    MappedData of this module (host-side list surgery; no device work): pairs that touch a problem interval, pairs with one
    mate mapped whose other mate exists, and all discordant / unmapped pairs go to md.needToRemap (in the Java's order:
    concordant, left, right, discordant, unmapped); the records that stay are moved onto assembly coordinates and
    reordered concordant, leftMateMapped, rightMateMapped — the order of the rebuilt mappedData.mappedReads.
    problem_intervals: (start, end) pairs; genome_to_assembly: int array; reads: the resident PairedReads."""
    rec = md.records
    iv = np.asarray(problem_intervals, np.int64).reshape(-1, 2)
    g2a = np.asarray(genome_to_assembly, np.int64)

    def touches(m):
        s = int(m["start"])
        return any(s + int(m["start2"]) < e and s + int(m["end2"]) > b for b, e in iv)

    def exists(pi, mate):
        off, ln = (reads.off1, reads.len1) if mate == 0 else (reads.off2, reads.len2)
        return not (ln[pi] == 1 and reads.bases[off[pi]] == ord("*"))

    need, kc, kl, kr = [], [], [], []
    for r in md.concordant:
        (need if touches(rec["m"][r, 0]) or touches(rec["m"][r, 1]) else kc).append(r)
    for r in md.leftMateMapped:
        if exists(md.pair_index[r], 1) or touches(rec["m"][r, 0]):
            need.append(r)
        else:
            kl.append(r)
    for r in md.rightMateMapped:
        if exists(md.pair_index[r], 0) or touches(rec["m"][r, 1]):
            need.append(r)
        else:
            kr.append(r)
    need += list(md.discordant) + list(md.unmapped)
    keep = np.array(kc + kl + kr, np.int64)
    out = rec[keep].copy()
    for r in range(len(out)):
        for k in range(2):
            m = out["m"][r, k]
            if m["present"]:  # adjustCoord (:662-667)
                s, s2, e2 = int(m["start"]), int(m["start2"]), int(m["end2"])
                out["m"]["end"][r, k] = g2a[s + e2 - 1] + 1
                out["m"]["start"][r, k] = g2a[s + s2]
                out["m"]["end2"][r, k] = e2 - s2
                out["m"]["start2"][r, k] = 0
    need_idx = md.pair_index[np.array(need, np.int64)] if need else np.zeros(0, np.int64)
    n_keep = len(keep)
    # the remapped pairs are parked behind the kept rows as unmapped placeholders so that needToRemap can name them
    park = np.zeros(len(need), PAIR_DTYPE)
    park["cls"] = _lib.CLS_UNMAPPED
    park["m"]["seq1_offset"] = -1
    md.records = np.concatenate([out, park])
    md.pair_index = np.concatenate([md.pair_index[keep], need_idx])
    md.needToRemap = np.arange(n_keep, n_keep + len(need), dtype=np.int64)
    return md


class MapperMSA:
    """MapperMSA (MapperMSA.java): seeding + pairing against the reference MSA index, no alignment."""

    def __init__(self, K=7, indel_tolerance=1.5, cutoffs=None, insert_len=1000, device=0, ctx=None):
        self.ctx = ctx or Context(K, 1.0 / indel_tolerance, indel_tolerance, cutoffs, insert_len, device=device)
        self.K = K
        self.low = np.float32(1.0) / np.float32(indel_tolerance)  # KMerCounterMSA.java:22-23
        self.high = np.float32(indel_tolerance)
        self.cutoffs = np.ascontiguousarray(cutoffs if cutoffs is not None else np.zeros(1024, np.int32), np.int32)
        self._reads = None
        self._aln = None

    def mapAllReads(self, reads, refAlignment):
        if reads is not self._reads:
            self.ctx.upload_reads(reads)
            self._reads = reads
        if refAlignment is not self._aln:
            self.ctx.set_msa(refAlignment, float(self.low), float(self.high), self.cutoffs)
            self._aln = refAlignment
        return self.ctx.map_pairs_msa()


class ConsensusBuilderSimple:
    """ConsensusBuilderSimple (ConsensusBuilderSimple.java): pileup of the last mapping + per-position argmax."""

    def __init__(self, mapper: Mapper, coverage_threshold=0, allreduce=None):
        self.mapper = mapper
        self.coverageThreshold = coverage_threshold
        self.mappedData = None
        self._allreduce = allreduce  # callable(dev_ptr, n_int32) summing the counts over ranks (multi-GPU)

    def buildConsensus(self, genome, reads):
        self.mappedData = self.mapper.mapReads(reads, genome)
        return self.getConsensusSimple(False, self.coverageThreshold, genome.seqB)

    def getConsensusSimple(self, saveUncovered, coverageThreshold, genome):
        ctx = self.mapper.ctx
        ctx.pileup(False)  # a MultiContext all-reduces inside
        if self._allreduce is not None:
            p, n = ctx.counts_device()
            self._allreduce(p, n)
        elif isinstance(ctx, Context):
            ctx.pileup_allreduce()  # one process per GPU: no-op unless comm_init_rank was called
        return ctx.consensus(genome, saveUncovered, coverageThreshold)
